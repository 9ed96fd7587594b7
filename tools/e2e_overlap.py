"""How well do a forward and an inverse host-buffer call overlap on one context? (PCIe full duplex)"""
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import jwave_pro_b200 as jw  # noqa: E402

B, n, J = 256, 65536, 6
w = jw.wavelets.Daubechies4()
hx = torch.rand((B, n), dtype=torch.float64).pin_memory()
hc = torch.empty((B, J + 1, n), dtype=torch.float64).pin_memory()
hc2 = torch.empty((B, J + 1, n), dtype=torch.float64).pin_memory()
hr = torch.empty((B, n), dtype=torch.float64).pin_memory()
X, C, C2, R = hx.numpy(), hc.numpy(), hc2.numpy(), hr.numpy()

for chunk, nbuf in ((128, 2), (128, 3), (128, 4), (64, 3), (64, 4), (256, 2), (256, 3)):
    ctx = jw.Context([0])
    ctx.set_tuning("h2d_chunk_mb", chunk)
    ctx.set_tuning("h2d_buffers", nbuf)
    t = jw.CudaMODWTTransform(w, context=ctx)
    t.forwardMODWTBatch(X, J, out=C)
    t.forwardMODWTBatch(X, J, out=C2)
    t.inverseMODWTBatch(C, out=R)

    def timed(fn, reps=4):
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e3

    f_ms = timed(lambda: t.forwardMODWTBatch(X, J, out=C))
    i_ms = timed(lambda: t.inverseMODWTBatch(C, out=R))

    def both():
        a = threading.Thread(target=lambda: t.forwardMODWTBatch(X, J, out=C2))
        b = threading.Thread(target=lambda: t.inverseMODWTBatch(C, out=R))
        a.start(), b.start()
        a.join(), b.join()

    both()
    b_ms = timed(both)
    print("nbuf %d chunk %3d MB: forward %.2f ms  inverse %.2f ms  both concurrently %.2f ms" % (nbuf, chunk, f_ms, i_ms, b_ms))
    ctx.close()
