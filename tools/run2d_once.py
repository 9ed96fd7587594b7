"""One 2-D forward + reverse on device buffers (for ncu launch lists): python tools/run2d_once.py fwt Haar1 32 4096 4096"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jwave_pro_b200 as jw  # noqa: E402

kind, cls, batch, rows, cols = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
lm = int(sys.argv[6]) if len(sys.argv) > 6 else int(math.log2(rows))
ln = int(sys.argv[7]) if len(sys.argv) > 7 else int(math.log2(cols))
t = (jw.CudaFastWaveletTransform if kind == "fwt" else jw.CudaWaveletPacketTransform)(jw.wavelets.create(cls))
x = torch.rand((batch, rows, cols), dtype=torch.float64, device="cuda")
c, r = torch.empty_like(x), torch.empty_like(x)
for _ in range(2):
    t.forward2DDevice(x.data_ptr(), c.data_ptr(), batch, rows, cols, lm, ln)
    t.reverse2DDevice(c.data_ptr(), r.data_ptr(), batch, rows, cols, lm, ln)
torch.cuda.synchronize()
print("pr", float((r - x).abs().max()))
