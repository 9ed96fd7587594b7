"""MODWT on lengths that are not powers of two, long filter / deep transform (the shapes whose deeper passes walk the
gcd(2^j0, n) cycles of the circular signal): device-resident forward + inverse, CUDA events, fused path against the
per-level generic kernels (what these shapes ran on before)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jwave_pro_b200 as jw

ctx = jw.default_context()
for cls, n, J, batch in (("Daubechies20", 99999, 8, 2048), ("Daubechies20", 100000, 8, 2048), ("Daubechies8", 65538, 9, 2048),
                         ("Daubechies4", 99999, 10, 2048)):
    w = jw.wavelets.create(cls)
    t = jw.CudaMODWTTransform(w, context=ctx)
    x = torch.rand((batch, n), dtype=torch.float64, device="cuda") * 2 - 1
    c = torch.empty((batch, (J + 1) * n), dtype=torch.float64, device="cuda")
    xr = torch.empty_like(x)
    st = torch.cuda.current_stream()
    for flags, name in ((0, "fused"), (jw.FLAG_FORCE_GENERIC, "generic")):
        def step():
            t.forwardMODWTDevice(x.data_ptr(), c.data_ptr(), batch, n, J, stream=st.cuda_stream, flags=flags)
            t.inverseMODWTDevice(c.data_ptr(), xr.data_ptr(), batch, n, J, stream=st.cuda_stream, flags=flags)
        for _ in range(3):
            step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(st)
        for _ in range(5):
            step()
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        err = float((xr - x).abs().max())
        print("%-13s n=%-6d J=%-2d batch=%d %-7s forward+inverse %.3f ms  %.1f Gsamples/s  round trip %.2e" % (
            cls, n, J, batch, name, ms, 2.0 * batch * n / ms / 1e6, err))
