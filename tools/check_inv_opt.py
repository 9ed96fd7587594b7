"""Inverse MODWT: outputs of every modwt_inv_opt variant against the default kernel path (bitwise) on a few shapes."""
import sys
import numpy as np
import jwave_pro_b200 as jw

ctx = jw.default_context()
rng = np.random.default_rng(1)
w = jw.wavelets.Daubechies4()
t = jw.CudaMODWTTransform(w)
bad = 0
for (b, n, J) in ((64, 65536, 6), (16, 100000, 6), (8, 4096, 5), (32, 65536, 4), (4, 1 << 20, 8)):
    x = rng.uniform(-1, 1, size=(b, n))
    ctx.set_tuning("modwt_inv_opt", 0)
    c = t.forwardMODWTBatch(x, J)
    ref = t.inverseMODWTBatch(c)
    for opt in (1, 2, 3, 7):
        ctx.set_tuning("modwt_inv_opt", opt)
        got = t.inverseMODWTBatch(c)
        same = np.array_equal(got, ref)
        err = float(np.max(np.abs(got - x)))
        print("b=%d n=%d J=%d opt=%d bitwise_equal=%s pr_err=%.2e" % (b, n, J, opt, same, err))
        bad += (not same) or err > 1e-10
ctx.set_tuning("modwt_inv_opt", 0)
sys.exit(1 if bad else 0)
