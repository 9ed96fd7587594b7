"""One sliding-window forward MODWT on device buffers (for ncu): python tools/run_windows_once.py Daubechies4 16777216 512 64 8"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jwave_pro_b200 as jw  # noqa: E402

cls, total, window, hop, J = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
ctx = jw.default_context()
t = jw.CudaMODWTTransform(jw.wavelets.create(cls))
g, h = (np.ascontiguousarray(v) for v in t._filters())
nwin = (total - window) // hop + 1
x = torch.rand(total, dtype=torch.float64, device="cuda")
c = torch.empty((nwin, J + 1, window), dtype=torch.float64, device="cuda")
dp = ctypes.POINTER(ctypes.c_double)
lib = jw._native.load()
for _ in range(2):
    rc = lib.jwc_modwt_forward_windows_dev(ctx.handle, 0, ctypes.c_void_p(1), ctypes.c_void_p(x.data_ptr()),
                                           ctypes.c_void_p(c.data_ptr()), total, window, hop, J, g.ctypes.data_as(dp),
                                           h.ctypes.data_as(dp), len(g), 0)
    assert rc == 0
torch.cuda.synchronize()
print("ok", nwin)
