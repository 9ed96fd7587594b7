#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "3d" 2>&1 | tail -15 > gpurun_out/r4k_pytest.txt; cat gpurun_out/r4k_pytest.txt
python - <<'PY'
import torch, time, sys, os
sys.path.insert(0, os.getcwd())
import jwave_pro_b200 as jw
# 3-D FWT timing: 64 cubes of 256^3 (8 GiB in, 8 GiB out), Daubechies4 full depth
t = jw.CudaFastWaveletTransform(jw.wavelets.Daubechies4())
B, n = 16, 256
x = torch.rand((B, n, n, n), dtype=torch.float64, device="cuda") * 2 - 1
c = torch.empty_like(x); xr = torch.empty_like(x)
st = torch.cuda.current_stream()
for name, L in (("full depth", 8), ("3 levels", 3)):
    def step():
        t.forward3DDevice(x.data_ptr(), c.data_ptr(), B, n, n, n, L, L, L, stream=st.cuda_stream)
        t.reverse3DDevice(c.data_ptr(), xr.data_ptr(), B, n, n, n, L, L, L, stream=st.cuda_stream)
    for _ in range(3): step()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    torch.cuda.synchronize(); e0.record(st)
    for _ in range(5): t.forward3DDevice(x.data_ptr(), c.data_ptr(), B, n, n, n, L, L, L, stream=st.cuda_stream)
    e1.record(st)
    for _ in range(5): t.reverse3DDevice(c.data_ptr(), xr.data_ptr(), B, n, n, n, L, L, L, stream=st.cuda_stream)
    e2.record(st); torch.cuda.synchronize()
    f, r = e0.elapsed_time(e1) / 5, e1.elapsed_time(e2) / 5
    gb = 48.0 * B * n ** 3 / 1e9    # three passes (rows, columns, first axis), each one read + one write
    print("fwt3d Daubechies4 %d x %d^3 %s: forward %.3f ms (%.0f GB/s of 48 B/sample), inverse %.3f ms (%.0f GB/s), round trip %.2e" % (
        B, n, name, f, gb / f * 1e3, r, gb / r * 1e3, float((xr - x).abs().max())))
PY
