"""Device-resident timing of the 2-D FWT / WPT (CUDA events on the launch stream, inputs larger than L2).

Algorithmic bytes per sample = 32: the row pass and the column pass each read and write the matrix once (a full-depth
separable 2-D transform of a matrix much larger than shared memory cannot do with fewer passes).  Prints forward and
inverse time, achieved GB/s and the fraction of the measured HBM peak, and the row pass alone beside it.
"""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jwave_pro_b200 as jw  # noqa: E402

CASES = [("fwt", "Haar1", 32, 4096, 4096, None, None), ("fwt", "Daubechies4", 32, 4096, 4096, None, None),
         ("fwt", "Daubechies8", 32, 4096, 4096, None, None), ("fwt", "Daubechies4", 128, 2048, 2048, 3, 3),
         ("wpt", "Symlet8", 32, 4096, 4096, 3, 3), ("fwt", "Daubechies20", 32, 4096, 4096, None, None),
         ("wpt", "Daubechies4", 32, 4096, 4096, 3, 3), ("wpt", "Haar1", 32, 4096, 4096, 6, 6)]


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
    only = os.environ.get("JWC_CASES")
    st = torch.cuda.Stream()
    for ci, (kind, cls, batch, rows, cols, lm, ln) in enumerate(CASES):
        if only and str(ci) not in only.split(","):
            continue
        w = jw.wavelets.create(cls)
        t = (jw.CudaFastWaveletTransform if kind == "fwt" else jw.CudaWaveletPacketTransform)(w, context=ctx)
        lm = int(math.log2(rows)) if lm is None else lm
        ln = int(math.log2(cols)) if ln is None else ln
        x = torch.rand((batch, rows, cols), dtype=torch.float64, device="cuda") * 2 - 1
        c = torch.empty_like(x)
        r = torch.empty_like(x)

        def timed(fn, reps=5):
            for _ in range(10):   # also ramps the clocks for the first case
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(reps):
                fn()
            e1.record(st)
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        f_ms = timed(lambda: t.forward2DDevice(x.data_ptr(), c.data_ptr(), batch, rows, cols, lm, ln, stream=st.cuda_stream))
        i_ms = timed(lambda: t.reverse2DDevice(c.data_ptr(), r.data_ptr(), batch, rows, cols, lm, ln, stream=st.cuda_stream))
        # the first timed loop of a new shape also pays one-time costs (module load of new kernel instantiations, pool
        # growth while the host runs ahead): time the forward again and keep the steady-state figure
        f_ms = min(f_ms, timed(lambda: t.forward2DDevice(x.data_ptr(), c.data_ptr(), batch, rows, cols, lm, ln, stream=st.cuda_stream)))
        row_ms = timed(lambda: t.forwardDevice(x.data_ptr(), c.data_ptr(), batch * rows, cols, ln, stream=st.cuda_stream))
        t.forward2DDevice(x.data_ptr(), c.data_ptr(), batch, rows, cols, lm, ln, stream=st.cuda_stream)
        t.reverse2DDevice(c.data_ptr(), r.data_ptr(), batch, rows, cols, lm, ln, stream=st.cuda_stream)
        torch.cuda.synchronize()
        pr = float((r - x).abs().max())
        n = batch * rows * cols
        gbs = lambda ms: 32.0 * n / (ms * 1e-3) / 1e9  # noqa: E731
        print("%s2d %-12s %3d x %d x %d lvl %d/%d: fwd %.3f ms (%.0f GB/s, %.2f)  inv %.3f ms (%.0f GB/s, %.2f)  "
              "row pass alone %.3f ms  pr=%.1e" % (kind, cls, batch, rows, cols, lm, ln, f_ms, gbs(f_ms), gbs(f_ms) / peak,
                                                   i_ms, gbs(i_ms), gbs(i_ms) / peak, row_ms, pr), flush=True)
        del x, c, r
    ctx.close()


if __name__ == "__main__":
    main()
