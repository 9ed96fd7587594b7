"""Device-resident timing of the sliding-window forward MODWT (the reference's MODWTSlidingWindowTest shape, scaled up).

Algorithmic bytes per window: 8*hop read (each series sample is new to the device once) + 8*(J+1)*window written.
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jwave_pro_b200 as jw  # noqa: E402

CASES = [("Haar1", 1 << 24, 512, 64, 8), ("Daubechies4", 1 << 24, 512, 64, 8), ("Daubechies4", 1 << 26, 4096, 1024, 6),
         ("Daubechies4", 1 << 24, 512, 64, 3)]


def main():
    peak = 6545.3
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass
    ctx = jw.Context([0])
    for kv in filter(None, os.environ.get("JWC_TUNE", "").split(",")):
        k, v = kv.split("=")
        ctx.set_tuning(k, int(v))
    lib = jw._native.load()
    dp = ctypes.POINTER(ctypes.c_double)
    st = torch.cuda.Stream()
    for cls, total, window, hop, J in CASES:
        t = jw.CudaMODWTTransform(jw.wavelets.create(cls), context=ctx)
        g, h = (np.ascontiguousarray(v) for v in t._filters())
        nwin = (total - window) // hop + 1
        x = torch.rand(total, dtype=torch.float64, device="cuda") * 2 - 1
        c = torch.empty((nwin, J + 1, window), dtype=torch.float64, device="cuda")

        def fn():
            rc = lib.jwc_modwt_forward_windows_dev(ctx.handle, 0, ctypes.c_void_p(st.cuda_stream),
                                                   ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(c.data_ptr()), total,
                                                   window, hop, J, g.ctypes.data_as(dp), h.ctypes.data_as(dp), len(g), 0)
            assert rc == 0, lib.jwc_last_error()

        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(5):
            fn()
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        byts = 8.0 * (total + nwin * (J + 1) * window)
        print("windows %-12s series %d window %d hop %d J %d: %d windows, %.3f ms, %.1f Mwindows/s, %.0f GB/s (%.2f of HBM)"
              % (cls, total, window, hop, J, nwin, ms, nwin / ms / 1e3, byts / ms / 1e6, byts / ms / 1e6 / peak), flush=True)
        del x, c
    ctx.close()


if __name__ == "__main__":
    main()
