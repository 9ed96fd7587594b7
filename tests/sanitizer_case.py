"""Small driver for compute-sanitizer (memcheck / racecheck): every fused kernel family once, small shapes.
    compute-sanitizer --tool memcheck python tests/sanitizer_case.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jwave_pro_b200 as jw  # noqa: E402

rng = np.random.default_rng(0)
for cls, n, J in (("Daubechies4", 8192, 6), ("Daubechies20", 16384, 8), ("Haar1", 4096, 5)):
    w = jw.wavelets.create(cls)
    x = rng.uniform(-1, 1, size=(3, n))
    t = jw.CudaMODWTTransform(w)
    c = t.forwardMODWTBatch(x, J)
    assert np.max(np.abs(t.inverseMODWTBatch(c) - x)) < 1e-9
    for T in (jw.CudaFastWaveletTransform, jw.CudaWaveletPacketTransform):
        tr = T(w)
        lv = 7
        y = tr.forwardBatch(x, lv)
        assert np.max(np.abs(tr.reverseBatch(y, lv) - x)) < 1e-9
# awkward shapes: generic fallbacks, scalar loaders
w = jw.wavelets.Symlet8()
t = jw.CudaMODWTTransform(w)
for n, J in ((100, 6), (65537, 4), (8, 3)):
    x = rng.uniform(-1, 1, size=(2, n))
    assert np.max(np.abs(t.inverseMODWTBatch(t.forwardMODWTBatch(x, J)) - x)) < 1e-9
f = jw.CudaFastWaveletTransform(jw.wavelets.Daubechies20())
for n in (2, 4, 64):
    x = rng.uniform(-1, 1, size=(2, n))
    p = int(np.log2(n))
    assert np.max(np.abs(f.reverseBatch(f.forwardBatch(x, p), p) - x)) < 1e-8
print("sanitizer case ok")
