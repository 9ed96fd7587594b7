"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle and the reference's KATs.

Bars (BASELINE.json north_star): max|err| <= 1e-12 * max|x| for the default (FMA) kernels, perfect reconstruction
<= 1e-10; in JWC_FLAG_EXACT mode (unfused mul/add in the reference's order) the result must equal the oracle
bit for bit.  Nothing here reads /root/reference.
"""
import ctypes
import math
import os
import threading

import numpy as np
import pytest

from conftest import chirp, splitmix_uniform

pytestmark = pytest.mark.gpu

TOL = 1e-12   # north_star parity tolerance, relative to max|x|
PR_TOL = 1e-10


def _maxerr(a, b, x):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)))) / max(float(np.max(np.abs(x))), 1e-300)


def _modwt_oracle(oracle, w, X, J, nthreads=8):
    g, h = oracle.modwt_filters(w.getScalingDeComposition(), w.getWaveletDeComposition())
    return oracle.batch("modwt_fwd", X, J, g, h, nthreads=nthreads), (g, h)


# ----------------------------------------------------------------------------------------------------------------
# reference KATs on the GPU
# ----------------------------------------------------------------------------------------------------------------

def test_kat_modwt_haar(jw, gpu_ctx, kats):
    k = kats["modwt_haar_level1"]
    t = jw.CudaMODWTTransform(jw.wavelets.Haar1())
    for flags in (0, jw.FLAG_EXACT, jw.FLAG_FORCE_GENERIC):
        c = t.forwardMODWT(k["input"], 1, flags=flags)
        np.testing.assert_allclose(c[0], k["D1"], atol=1e-9)
        np.testing.assert_allclose(c[1], k["A1"], atol=1e-9)
        np.testing.assert_allclose(t.inverseMODWT(c, flags=flags), k["input"], atol=1e-9)


def test_kat_haar_fwt_fixture(jw, gpu_ctx, kats):
    k = kats["haar_fwt_level1"]
    out = jw.CudaFastWaveletTransform(jw.wavelets.Haar1()).forward(k["input"], 1)
    np.testing.assert_allclose(out[:4], k["approx"], atol=1e-10)
    np.testing.assert_allclose(out[4:], k["detail"], atol=1e-10)


@pytest.mark.parametrize("n", [4, 64])
def test_kat_all_ones_ladders(jw, gpu_ctx, n):
    """SteppingTest.java:37-314 for the 44 in-scope wavelets: includes filters far longer than the signal (L=40, n=4)."""
    ones = np.ones(n)
    for w in jw.wavelets.create2arr():
        for T in (jw.CudaFastWaveletTransform, jw.CudaWaveletPacketTransform):
            t = T(w)
            for p in range(int(math.log2(n)) + 1):
                e = np.zeros(n)
                e[: n >> p] = 2.0 ** (p / 2.0)
                c = t.forward(ones, p)
                np.testing.assert_allclose(c, e, atol=1e-8, err_msg="%s %s level %d" % (T.__name__, w.getName(), p))
                np.testing.assert_allclose(t.reverse(c, p), ones, atol=1e-8)
            d = t.decompose(ones)
            np.testing.assert_allclose(t.recompose(d, int(math.log2(n))), ones, atol=1e-8)


# ----------------------------------------------------------------------------------------------------------------
# exact mode: bit-identical to the oracle
# ----------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("cls,n,J", [("Haar1", 8, 3), ("Symlet8", 8, 3), ("Daubechies4", 1024, 3),
                                     ("Daubechies4", 100, 6), ("Daubechies6", 288, 5), ("Daubechies20", 1000, 9),
                                     ("Coiflet5", 500, 8), ("Daubechies20", 64, 6), ("Symlet20", 4096, 12)])
def test_exact_modwt_bitwise(jw, gpu_ctx, oracle, cls, n, J):
    w = jw.wavelets.create(cls)
    t = jw.CudaMODWTTransform(w)
    X = splitmix_uniform(n * 31 + J, (3, n))
    ref, (g, h) = _modwt_oracle(oracle, w, X, J)
    got = t.forwardMODWTBatch(X, J, flags=jw.FLAG_EXACT)
    assert np.array_equal(got, ref)
    back = t.inverseMODWTBatch(got, flags=jw.FLAG_EXACT)
    assert np.array_equal(back, oracle.batch("modwt_inv", ref, J, g, h, nthreads=4))
    # (no absolute PR bar here: e.g. Coiflet5's table reconstructs only to ~3e-8 in the reference itself)


@pytest.mark.parametrize("cls", ["Haar1", "Daubechies2", "Daubechies4", "Daubechies8", "Daubechies20", "Symlet8",
                                 "Coiflet1", "Coiflet5"])
@pytest.mark.parametrize("n", [2, 4, 16, 64, 1024])
def test_exact_fwt_wpt_bitwise(jw, gpu_ctx, oracle, cls, n):
    w = jw.wavelets.create(cls)
    s, wv = w.getScalingDeComposition(), w.getWaveletDeComposition()
    X = splitmix_uniform(n + len(s), (2, n))
    p = int(math.log2(n))
    for lvl in sorted({0, 1, p // 2, p}):
        for T, fo, ro in ((jw.CudaFastWaveletTransform, "fwt_fwd", "fwt_rev"),
                          (jw.CudaWaveletPacketTransform, "wpt_fwd", "wpt_rev")):
            t = T(w)
            ref = oracle.batch(fo, X, lvl, s, wv)
            got = t.forwardBatch(X, lvl, flags=jw.FLAG_EXACT)
            assert np.array_equal(got, ref), (T.__name__, cls, n, lvl)
            rref = oracle.batch(ro, ref, lvl, s, wv)
            rgot = t.reverseBatch(got, lvl, flags=jw.FLAG_EXACT)
            assert np.array_equal(rgot, rref), (T.__name__, cls, n, lvl, "reverse")


# ----------------------------------------------------------------------------------------------------------------
# default kernels vs the oracle at the BASELINE configurations (a subset of signals at the full length)
# ----------------------------------------------------------------------------------------------------------------

def _inputs(seed, batch, n):
    X = splitmix_uniform(seed, (batch, n))
    X[batch // 2:] = chirp(batch - batch // 2, n)   # random + chirp, SURVEY.md section 8d
    return X


@pytest.mark.parametrize("cls,n,J,batch", [
    ("Daubechies4", 1024, 3, 1),        # C1 (the reference's own CPU-runnable case)
    ("Daubechies4", 65536, 6, 6),       # C2 shape
    ("Daubechies20", 65536, 8, 4),      # C5 shape
    ("Symlet8", 65536, 6, 3),
    ("Haar1", 65536, 13, 2),            # deepest level the reference allows
    ("Daubechies4", 8192, 13, 2),
    ("Daubechies20", 8192, 13, 1),      # (L-1)*2^12 = 159744 > n: the filter wraps many times
    ("Daubechies8", 4096, 5, 37),       # ragged batch
    ("Daubechies4", 100, 6, 5),         # not a power of two (MODWTInverseTest.java:20-92)
    ("Daubechies6", 1000, 9, 3),
    ("Daubechies4", 65538, 6, 2),       # even but not a multiple of anything convenient
    ("Daubechies4", 40000, 7, 3),
    ("Daubechies4", 65537, 4, 2),       # odd length
    ("Symlet8", 8, 3, 2),               # filter longer than the signal (MODWTFFTConvolutionTest.java:42-56)
    # 2^j0 does not divide n: the deeper fused passes walk the gcd(2^j0, n) interleaved cycles of the circular signal
    ("Daubechies20", 99999, 8, 2),      # odd: one cycle, single-phase passes
    ("Daubechies20", 100000, 8, 2),     # 2^5 | n: 32 cycles at j0 = 6
    ("Daubechies8", 65538, 9, 2),       # 2 x odd: two cycles, 16-byte rows
    ("Daubechies4", 3000, 11, 2),       # 8 cycles of 375 positions, deep passes with halos longer than a cycle
    ("Haar1", 12345, 10, 3),
])
def test_modwt_matches_oracle(jw, gpu_ctx, oracle, cls, n, J, batch):
    w = jw.wavelets.create(cls)
    t = jw.CudaMODWTTransform(w)
    X = _inputs(n + J, batch, n)
    ref, (g, h) = _modwt_oracle(oracle, w, X, J)
    got = t.forwardMODWTBatch(X, J)
    assert _maxerr(got, ref, X) <= TOL
    gen = t.forwardMODWTBatch(X, J, flags=jw.FLAG_FORCE_GENERIC)
    assert _maxerr(gen, ref, X) <= TOL
    back = t.inverseMODWTBatch(ref)
    assert _maxerr(back, oracle.batch("modwt_inv", ref, J, g, h, nthreads=8), X) <= TOL
    assert _maxerr(t.inverseMODWTBatch(got), X, X) <= PR_TOL
    # single-signal API and flat 1-D API agree with the batch API
    one = t.forwardMODWT(X[0], J)
    assert np.array_equal(one, got[0])
    if jw.BasicTransform.isBinary(n):
        flat = t.forward(X[0], J)
        assert np.array_equal(flat, got[0].reshape(-1))
        assert _maxerr(t.reverse(flat, J), X[0], X[0]) <= PR_TOL


@pytest.mark.parametrize("kind", ["fwt", "wpt"])
@pytest.mark.parametrize("cls,n,lvl,batch", [
    ("Haar1", 1 << 20, 20, 2),          # C3
    ("Daubechies8", 1 << 20, 20, 2),    # C3
    ("Symlet8", 65536, 6, 4),           # C4
    ("Daubechies20", 65536, 16, 2),
    ("Daubechies4", 1 << 17, 9, 3),
    ("Coiflet3", 2048, 11, 9),
    ("Daubechies8", 4096, 3, 33),
    ("Daubechies20", 32, 5, 3),
    ("Daubechies4", 2, 1, 2),
    ("Haar1", 1, 0, 3),
])
def test_fwt_wpt_match_oracle(jw, gpu_ctx, oracle, kind, cls, n, lvl, batch):
    if kind == "wpt" and n >= (1 << 20):
        lvl = 8   # a 20-level packet tree of 2^20 samples costs the CPU oracle too long
    w = jw.wavelets.create(cls)
    s, wv = w.getScalingDeComposition(), w.getWaveletDeComposition()
    t = (jw.CudaFastWaveletTransform if kind == "fwt" else jw.CudaWaveletPacketTransform)(w)
    X = _inputs(n + lvl, batch, n)
    ref = oracle.batch(kind + "_fwd", X, lvl, s, wv, nthreads=8)
    got = t.forwardBatch(X, lvl)
    assert _maxerr(got, ref, X) <= TOL
    gen = t.forwardBatch(X, lvl, flags=jw.FLAG_FORCE_GENERIC)
    assert _maxerr(gen, ref, X) <= TOL
    rref = oracle.batch(kind + "_rev", ref, lvl, s, wv, nthreads=8)
    assert _maxerr(t.reverseBatch(ref, lvl), rref, X) <= TOL
    assert _maxerr(t.reverseBatch(got, lvl), X, X) <= PR_TOL
    assert np.array_equal(t.forward(X[0], lvl), got[0])
    # every intermediate level is reachable, like the reference's stepping API
    for l2 in sorted({0, 1, lvl // 2}):
        if l2 <= lvl:
            r2 = oracle.batch(kind + "_fwd", X[:1], l2, s, wv)
            assert _maxerr(t.forwardBatch(X[:1], l2), r2, X) <= TOL


# ----------------------------------------------------------------------------------------------------------------
# host pipeline, threads, degenerate batches
# ----------------------------------------------------------------------------------------------------------------

def test_host_pipeline_many_chunks(jw, oracle):
    ctx = jw.Context()
    ctx.set_tuning("h2d_chunk_mb", 1)   # forces many double-buffered chunks
    w = jw.wavelets.Daubechies4()
    t = jw.CudaMODWTTransform(w, context=ctx)
    X = _inputs(77, 41, 8192)
    ref, (g, h) = _modwt_oracle(oracle, w, X, 5)
    got = t.forwardMODWTBatch(X, 5)
    assert _maxerr(got, ref, X) <= TOL
    assert _maxerr(t.inverseMODWTBatch(got), X, X) <= PR_TOL
    f = jw.CudaFastWaveletTransform(w, context=ctx)
    s, wv = w.getScalingDeComposition(), w.getWaveletDeComposition()
    assert _maxerr(f.forwardBatch(X, 13), oracle.batch("fwt_fwd", X, 13, s, wv, nthreads=8), X) <= TOL
    assert ctx.launch_count() > 0
    ctx.close()


def test_empty_batch_and_bad_arguments(jw, gpu_ctx):
    from jwave_pro_b200 import _native
    t = jw.CudaMODWTTransform(jw.wavelets.Haar1())
    out = t.forwardMODWTBatch(np.empty((0, 64)), 3)
    assert out.shape == (0, 4, 64)
    lib = _native.load()
    x = np.ones(8)
    o = np.empty(16)
    g = np.array([0.5, 0.5])
    dp = _native._dp
    rc = lib.jwc_modwt_forward(gpu_ctx.handle, x.ctypes.data, o.ctypes.data, 1, 8, 0, g.ctypes.data_as(dp),
                               g.ctypes.data_as(dp), 2, 0)
    assert rc == -1 and b"level" in lib.jwc_last_error()
    rc = lib.jwc_fwt_forward(gpu_ctx.handle, x.ctypes.data, o.ctypes.data, 1, 7, 1, g.ctypes.data_as(dp),
                             g.ctypes.data_as(dp), 2, 0)
    assert rc == -1 and b"2^p" in lib.jwc_last_error()
    rc = lib.jwc_fwt_forward(gpu_ctx.handle, x.ctypes.data, o.ctypes.data, 1, 8, 4, g.ctypes.data_as(dp),
                             g.ctypes.data_as(dp), 2, 0)
    assert rc == -1 and b"out of range" in lib.jwc_last_error()
    rc = lib.jwc_fwt_forward(gpu_ctx.handle, None, o.ctypes.data, 1, 8, 1, g.ctypes.data_as(dp),
                             g.ctypes.data_as(dp), 2, 0)
    assert rc == -1
    rc = lib.jwc_wpt_forward(gpu_ctx.handle, x.ctypes.data, o.ctypes.data, 1, 8, 1, g.ctypes.data_as(dp),
                             g.ctypes.data_as(dp), 65, 0)
    assert rc == -1 and b"filter length" in lib.jwc_last_error()


def test_concurrent_host_calls_use_separate_lanes(jw, oracle):
    """A forward and an inverse host-buffer call in flight at once on ONE context (the streaming pattern bench.py's
    e2e leg uses): each gets its own stream lane, results stay exact, many pipeline chunks each."""
    ctx = jw.Context([0])
    ctx.set_tuning("h2d_chunk_mb", 1)
    w = jw.wavelets.Daubechies4()
    t = jw.CudaMODWTTransform(w, context=ctx)
    X = splitmix_uniform(77, (96, 8192))
    ref, (g, h) = _modwt_oracle(oracle, w, X, 5)
    res = {}

    def fwd(i):
        res["f%d" % i] = t.forwardMODWTBatch(X, 5)

    def inv(i):
        res["i%d" % i] = t.inverseMODWTBatch(ref)

    th = [threading.Thread(target=f, args=(i,)) for i in range(3) for f in (fwd, inv)]
    [x.start() for x in th]
    [x.join() for x in th]
    back = oracle.batch("modwt_inv", ref, 5, g, h, nthreads=8)
    for i in range(3):
        assert _maxerr(res["f%d" % i], ref, X) <= TOL
        assert _maxerr(res["i%d" % i], back, X) <= TOL
        assert np.array_equal(res["f%d" % i], res["f0"]) and np.array_equal(res["i%d" % i], res["i0"])
    ctx.close()


def test_thread_safety_shared_instance(jw, gpu_ctx, oracle):
    """MODWTThreadSafetyTest.java:24-104: 10 threads x iterations on ONE instance, clearFilterCache every 10th."""
    w = jw.wavelets.Daubechies4()
    t = jw.CudaMODWTTransform(w)
    x = splitmix_uniform(9, (512,))
    g, h = oracle.modwt_filters(w.getScalingDeComposition(), w.getWaveletDeComposition())
    ref = oracle.modwt_forward(x, 4, g, h)
    errs = []

    def work(tid):
        try:
            for it in range(30):
                c = t.forwardMODWT(x, 4)
                if np.max(np.abs(c - ref)) > 1e-10:
                    errs.append((tid, it))
                if it % 10 == 0:
                    t.clearFilterCache()
        except Exception as e:   # noqa: BLE001
            errs.append((tid, repr(e)))

    th = [threading.Thread(target=work, args=(i,)) for i in range(10)]
    [x_.start() for x_ in th]
    [x_.join() for x_ in th]
    assert not errs, errs[:3]


# ----------------------------------------------------------------------------------------------------------------
# full BASELINE sizes, device-resident: size-independent properties (round trip, linearity, shift invariance)
# ----------------------------------------------------------------------------------------------------------------

def _torch_inputs(torch, batch, n, seed):
    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed)
    return torch.rand((batch, n), dtype=torch.float64, device="cuda", generator=gen) * 2.0 - 1.0


@pytest.mark.parametrize("cls,n,J,batch", [("Daubechies4", 65536, 6, 4096), ("Daubechies20", 65536, 8, 1024)])
def test_modwt_full_size_properties(jw, gpu_ctx, oracle, cls, n, J, batch):
    import torch
    w = jw.wavelets.create(cls)
    t = jw.CudaMODWTTransform(w)
    st = torch.cuda.current_stream().cuda_stream
    x = _torch_inputs(torch, batch, n, 1234)
    c = torch.empty((batch, J + 1, n), dtype=torch.float64, device="cuda")
    t.forwardMODWTDevice(x.data_ptr(), c.data_ptr(), batch, n, J, stream=st)
    xr = torch.empty_like(x)
    t.inverseMODWTDevice(c.data_ptr(), xr.data_ptr(), batch, n, J, stream=st)
    torch.cuda.synchronize()
    assert float((xr - x).abs().max()) <= PR_TOL
    # energy: sum of squares of all coefficients == energy of the signal (MODWTTransformTest.java:74-89)
    e_x = (x * x).sum(dim=1)
    e_c = (c * c).sum(dim=(1, 2))
    assert float(((e_c - e_x).abs() / e_x).max()) < 1e-9
    # spot rows against the oracle: first / last / a few pseudo-random signals
    rows = sorted({0, 1, batch - 1, batch // 2, (batch * 7) // 13, 3 % batch})
    X = x[rows].cpu().numpy()
    ref, _ = _modwt_oracle(oracle, w, X, J)
    assert _maxerr(c[rows].cpu().numpy(), ref, X) <= TOL
    # shift invariance: MODWT(roll(x, s)) == roll(MODWT(x), s) (PropertyBasedTest.java:316-357)
    xs = torch.roll(x[:64], shifts=12345, dims=1).contiguous()
    cs = torch.empty((64, J + 1, n), dtype=torch.float64, device="cuda")
    t.forwardMODWTDevice(xs.data_ptr(), cs.data_ptr(), 64, n, J, stream=st)
    torch.cuda.synchronize()
    assert float((cs - torch.roll(c[:64], shifts=12345, dims=2)).abs().max()) <= 1e-12
    # linearity: T(a x + b y) == a T(x) + b T(y)
    y = _torch_inputs(torch, 64, n, 99)
    cy = torch.empty_like(cs)
    t.forwardMODWTDevice(y.data_ptr(), cy.data_ptr(), 64, n, J, stream=st)
    z = (0.75 * x[:64] - 1.5 * y).contiguous()
    cz = torch.empty_like(cs)
    t.forwardMODWTDevice(z.data_ptr(), cz.data_ptr(), 64, n, J, stream=st)
    torch.cuda.synchronize()
    assert float((cz - (0.75 * c[:64] - 1.5 * cy)).abs().max()) <= 1e-12


@pytest.mark.parametrize("kind,cls,n,lvl,batch", [("fwt", "Haar1", 1 << 20, 20, 1024),
                                                   ("fwt", "Daubechies8", 1 << 20, 20, 1024),
                                                   ("wpt", "Symlet8", 65536, 6, 512)])
def test_fwt_wpt_full_size_properties(jw, gpu_ctx, oracle, kind, cls, n, lvl, batch):
    import torch
    w = jw.wavelets.create(cls)
    s, wv = w.getScalingDeComposition(), w.getWaveletDeComposition()
    t = (jw.CudaFastWaveletTransform if kind == "fwt" else jw.CudaWaveletPacketTransform)(w)
    st = torch.cuda.current_stream().cuda_stream
    x = _torch_inputs(torch, batch, n, 4321)
    c = torch.empty_like(x)
    xr = torch.empty_like(x)
    t.forwardDevice(x.data_ptr(), c.data_ptr(), batch, n, lvl, stream=st)
    t.reverseDevice(c.data_ptr(), xr.data_ptr(), batch, n, lvl, stream=st)
    torch.cuda.synchronize()
    assert float((xr - x).abs().max()) <= PR_TOL
    e_x = (x * x).sum(dim=1)
    e_c = (c * c).sum(dim=1)
    assert float(((e_c - e_x).abs() / e_x).max()) < 1e-9   # orthonormal transform (PropertyBasedTest.java:138-202)
    rows = sorted({0, batch - 1, batch // 3})
    X = x[rows].cpu().numpy()
    ref = oracle.batch(kind + "_fwd", X, lvl, s, wv, nthreads=8)
    assert _maxerr(c[rows].cpu().numpy(), ref, X) <= TOL
    y = _torch_inputs(torch, 8, n, 5)
    cy = torch.empty_like(y)
    t.forwardDevice(y.data_ptr(), cy.data_ptr(), 8, n, lvl, stream=st)
    z = (2.0 * x[:8] + 0.5 * y).contiguous()
    cz = torch.empty_like(y)
    t.forwardDevice(z.data_ptr(), cz.data_ptr(), 8, n, lvl, stream=st)
    torch.cuda.synchronize()
    assert float((cz - (2.0 * c[:8] + 0.5 * cy)).abs().max()) <= 1e-11


def test_multi_device_context_shards_by_signal(jw, oracle):
    """jwc_create(devices...) fans a host batch out over the devices (contiguous blocks of signals, no collective)."""
    import torch
    ndev = torch.cuda.device_count()
    # four device slots: distinct GPUs when the box has them, otherwise several slots on the same GPU (each slot has its
    # own streams, staging and host thread, so the fan-out / shard arithmetic of the C layer is exercised either way)
    devs = [i % ndev for i in range(4)]
    ctx = jw.Context(devs)
    assert ctx.num_devices() == 4
    w = jw.wavelets.Daubechies4()
    t = jw.CudaMODWTTransform(w, context=ctx)
    X = _inputs(5, 37, 4096)
    ref, (g, h) = _modwt_oracle(oracle, w, X, 5)
    got = t.forwardMODWTBatch(X, 5)
    assert _maxerr(got, ref, X) <= TOL
    assert _maxerr(t.inverseMODWTBatch(got), X, X) <= PR_TOL
    f = jw.CudaFastWaveletTransform(w, context=ctx)
    s, wv = w.getScalingDeComposition(), w.getWaveletDeComposition()
    assert _maxerr(f.forwardBatch(X, 12), oracle.batch("fwt_fwd", X, 12, s, wv, nthreads=8), X) <= TOL
    ctx.close()


@pytest.mark.parametrize("family", ["Haar1", "Daubechies", "Symlet", "Coiflet"])
def test_every_wavelet_through_the_fused_kernels(jw, gpu_ctx, oracle, family):
    """Every filter length (every template instantiation of the fused kernels, L = 2..40) against the oracle:
    MODWT J=5, FWT and WPT 6 levels on 3 signals of 8192 samples; tolerance 1e-12 * max|x|."""
    n, batch = 8192, 3
    X = _inputs(len(family), batch, n)
    for cls in [c for c in jw.wavelets.ALL_CLASSES if c.startswith(family)]:
        w = jw.wavelets.create(cls)
        s, wv = w.getScalingDeComposition(), w.getWaveletDeComposition()
        t = jw.CudaMODWTTransform(w)
        ref, (g, h) = _modwt_oracle(oracle, w, X, 5)
        got = t.forwardMODWTBatch(X, 5)
        assert _maxerr(got, ref, X) <= TOL, cls
        assert _maxerr(t.inverseMODWTBatch(ref), oracle.batch("modwt_inv", ref, 5, g, h, nthreads=8), X) <= TOL, cls
        for T, kind in ((jw.CudaFastWaveletTransform, "fwt"), (jw.CudaWaveletPacketTransform, "wpt")):
            tr = T(w)
            r = oracle.batch(kind + "_fwd", X, 6, s, wv, nthreads=8)
            assert _maxerr(tr.forwardBatch(X, 6), r, X) <= TOL, (cls, kind)
            assert _maxerr(tr.reverseBatch(r, 6), oracle.batch(kind + "_rev", r, 6, s, wv, nthreads=8), X) <= TOL, (cls, kind)


@pytest.mark.parametrize("cls,n,J", [("Daubechies4", 1 << 18, 6), ("Daubechies20", 300000, 5), ("Haar1", 65536, 13)])
def test_modwt_single_series_split_over_devices(jw, oracle, cls, n, J):
    """one long series in contiguous chunks on the context's devices, one halo exchange
    (cudaMemcpyPeerAsync, ring) per transform; equals the unsplit transform bit for bit and the oracle to 1e-12."""
    import torch
    P = 4   # distinct GPUs when the box has them, otherwise four slots on the same GPU (same ring of peer copies)
    devs = [i % torch.cuda.device_count() for i in range(P)]
    ctx = jw.Context(devs)
    w = jw.wavelets.create(cls)
    t = jw.CudaMODWTTransform(w, context=ctx)
    x = _inputs(n + J, 2, n)[1]          # a chirp
    x = x + 0.25 * splitmix_uniform(17, (n,))
    bounds = [n * p // P for p in range(P + 1)]
    xs, cs, xr = [], [], []
    for p in range(P):
        ln = bounds[p + 1] - bounds[p]
        xs.append(torch.from_numpy(x[bounds[p]:bounds[p + 1]].copy()).to("cuda:%d" % devs[p]))
        cs.append(torch.empty((J + 1, ln), dtype=torch.float64, device="cuda:%d" % devs[p]))
        xr.append(torch.empty(ln, dtype=torch.float64, device="cuda:%d" % devs[p]))
    for p in range(P):
        torch.cuda.synchronize(devs[p])
    t.forwardMODWTSplitDevice([a.data_ptr() for a in xs], [c.data_ptr() for c in cs], n, J)
    got = np.concatenate([c.cpu().numpy() for c in cs], axis=1)
    whole = jw.CudaMODWTTransform(w).forwardMODWT(x, J)
    assert np.array_equal(got, whole)
    g, h = oracle.modwt_filters(w.getScalingDeComposition(), w.getWaveletDeComposition())
    ref = oracle.modwt_forward(x, J, g, h)
    assert _maxerr(got, ref, x) <= TOL
    t.inverseMODWTSplitDevice([c.data_ptr() for c in cs], [a.data_ptr() for a in xr], n, J)
    back = np.concatenate([a.cpu().numpy() for a in xr])
    assert np.array_equal(back, jw.CudaMODWTTransform(w).inverseMODWT(whole))
    assert _maxerr(back, x, x) <= PR_TOL
    ctx.close()


@pytest.mark.parametrize("kind,cls,n,lvl", [
    ("fwt", "Daubechies4", 1 << 18, 18), ("fwt", "Daubechies8", 1 << 20, 20), ("fwt", "Haar1", 1 << 16, 16),
    ("fwt", "Daubechies20", 1 << 16, 5), ("fwt", "Symlet8", 1 << 17, 3), ("fwt", "Daubechies4", 4096, 12),
    ("fwt", "Coiflet3", 1 << 15, 0),
    ("wpt", "Symlet8", 1 << 18, 6), ("wpt", "Haar1", 1 << 16, 4), ("wpt", "Daubechies20", 1 << 17, 3),
    ("wpt", "Daubechies4", 1 << 16, 0)])
def test_fwt_wpt_single_series_split_over_devices(jw, oracle, kind, cls, n, lvl):
    """SURVEY 8e row 2, decimated transforms: one long series in P contiguous chunks on the context's device slots, halo
    exchange with the ring neighbour per fused pass (cudaMemcpyPeerAsync), the small FWT remainder gathered on slot 0.
    The bands assembled from the per-device chunks equal the unsplit transform (bit for bit on the split levels) and the
    oracle to 1e-12; the split inverse gives the series back."""
    import torch
    P = 4   # distinct GPUs when the box has them, otherwise four slots on the same GPU
    devs = [i % torch.cuda.device_count() for i in range(P)]
    ctx = jw.Context(devs)
    w = jw.wavelets.create(cls)
    T = jw.CudaFastWaveletTransform if kind == "fwt" else jw.CudaWaveletPacketTransform
    t, whole_t = T(w, context=ctx), T(w)
    x = _inputs(n + lvl, 2, n)[1] + 0.25 * splitmix_uniform(23, (n,))
    ln = n // P
    xs = [torch.from_numpy(x[p * ln:(p + 1) * ln].copy()).to("cuda:%d" % devs[p]) for p in range(P)]
    ys = [torch.empty(ln, dtype=torch.float64, device="cuda:%d" % devs[p]) for p in range(P)]
    zs = [torch.empty(ln, dtype=torch.float64, device="cuda:%d" % devs[p]) for p in range(P)]
    for d in set(devs):
        torch.cuda.synchronize(d)
    t.forwardSplitDevice([a.data_ptr() for a in xs], [a.data_ptr() for a in ys], n, lvl)
    got = t.splitLayoutToGlobal([a.cpu().numpy() for a in ys], n, lvl)
    whole = whole_t.forward(x, lvl)
    s, wv = w.getScalingDeComposition(), w.getWaveletDeComposition()
    ref = oracle.batch(kind + "_fwd", x[None, :], lvl, s, wv)[0]
    assert _maxerr(got, ref, x) <= TOL
    ls = t.splitLevels(n, lvl)
    assert 0 <= ls <= lvl
    if kind == "wpt":
        assert np.array_equal(got, whole)
    else:   # the bands computed split are the same kernels on the same samples; the remainder ran on one device
        assert np.array_equal(got[n >> ls:], whole[n >> ls:])
        assert _maxerr(got, whole, x) <= TOL
    # layout helpers are inverses of each other
    back_chunks = t.globalToSplitLayout(got, P, lvl)
    assert all(np.array_equal(back_chunks[p], ys[p].cpu().numpy()) for p in range(P))
    t.reverseSplitDevice([a.data_ptr() for a in ys], [a.data_ptr() for a in zs], n, lvl)
    back = np.concatenate([a.cpu().numpy() for a in zs])
    assert _maxerr(back, x, x) <= PR_TOL
    assert _maxerr(back, whole_t.reverse(whole, lvl), x) <= TOL
    ctx.close()


@pytest.mark.parametrize("kind", ["fwt", "wpt"])
def test_2d_device_buffers_aligned_to_8_bytes_only(jw, gpu_ctx, kind):
    """Device pointers at an odd element offset (16-byte alignment lost): the fused column launches stage rows with
    16-byte cp.async, and the packet transform's ping-pong reads the caller's OUTPUT buffer from the second launch on,
    so both pointers decide whether the fused launches may run (round-1 advice: only the input was checked)."""
    import torch
    B, rows, cols, lm, ln = 2, 512, 256, 6, 4     # deep column pass: three or more column launches
    w = jw.wavelets.Symlet8()
    t = (jw.CudaFastWaveletTransform if kind == "fwt" else jw.CudaWaveletPacketTransform)(w)
    x = torch.from_numpy(splitmix_uniform(77, (B * rows * cols,))).cuda()
    ref = torch.empty_like(x)
    t.forward2DDevice(x.data_ptr(), ref.data_ptr(), B, rows, cols, lm, ln)
    back_ref = torch.empty_like(x)
    t.reverse2DDevice(ref.data_ptr(), back_ref.data_ptr(), B, rows, cols, lm, ln)
    for in_off, out_off in ((1, 0), (0, 1), (1, 1)):
        xi = torch.empty(x.numel() + 2, dtype=torch.float64, device="cuda")
        yo = torch.full((x.numel() + 2,), 7.0, dtype=torch.float64, device="cuda")
        xi[in_off:in_off + x.numel()] = x
        assert xi[in_off:].data_ptr() % 16 == (8 if in_off else 0)
        t.forward2DDevice(xi[in_off:].data_ptr(), yo[out_off:].data_ptr(), B, rows, cols, lm, ln)
        torch.cuda.synchronize()
        got = yo[out_off:out_off + x.numel()]
        assert float((got - ref).abs().max()) <= 1e-12
        zi = torch.empty(x.numel() + 2, dtype=torch.float64, device="cuda")
        zi[in_off:in_off + x.numel()] = ref
        zo = torch.empty(x.numel() + 2, dtype=torch.float64, device="cuda")
        t.reverse2DDevice(zi[in_off:].data_ptr(), zo[out_off:].data_ptr(), B, rows, cols, lm, ln)
        torch.cuda.synchronize()
        assert float((zo[out_off:out_off + x.numel()] - back_ref).abs().max()) <= 1e-12
        assert float((back_ref - x).abs().max()) <= 1e-10


def test_many_short_lived_streams_do_not_pile_up_workspace(jw, oracle):
    """Workspace arenas are keyed by (device, caller stream); callers that come with ever new streams made them grow
    without bound (round-1 advice).  Beyond 48 arenas the idle ones are returned to the driver: 80 streams, one
    multi-pass transform each (needs scratch), results stay right."""
    import torch
    ctx = jw.Context()
    w = jw.wavelets.Daubechies20()
    t = jw.CudaMODWTTransform(w, context=ctx)
    x = splitmix_uniform(99, (2, 8192))
    g, h = oracle.modwt_filters(w.getScalingDeComposition(), w.getWaveletDeComposition())
    ref = oracle.batch("modwt_fwd", x, 8, g, h)
    dx = torch.from_numpy(x).cuda()
    outs = []
    for i in range(80):
        st = torch.cuda.Stream()
        dc = torch.empty((2, 9, 8192), dtype=torch.float64, device="cuda")
        st.wait_stream(torch.cuda.current_stream())
        t.forwardMODWTDevice(dx.data_ptr(), dc.data_ptr(), 2, 8192, 8, stream=st.cuda_stream)
        outs.append((st, dc))
    for st, dc in outs[::13] + outs[-1:]:
        st.synchronize()
        assert _maxerr(dc.cpu().numpy(), ref, x) <= TOL
    ctx.close()


def test_diagnostic_rooflines(jw, gpu_ctx):
    """jwc_diag_dfma_tflops / jwc_diag_copy_gbs: the in-run denominators of bench.py's fp64_frac / copy ceiling."""
    tf = gpu_ctx.dfma_tflops()
    gb = gpu_ctx.copy_gbs(1 << 28)
    assert 10.0 < tf < 60.0, tf        # B200: ~31-34 TFLOP/s fp64 FMA
    assert 1000.0 < gb < 9000.0, gb    # B200: ~6 TB/s read + write


def test_split_wpt_declines_too_many_levels(jw):
    import torch
    devs = [i % torch.cuda.device_count() for i in range(2)]
    ctx = jw.Context(devs)
    t = jw.CudaWaveletPacketTransform(jw.wavelets.Haar1(), context=ctx)
    n = 1 << 14
    bufs = [torch.zeros(n // 2, dtype=torch.float64, device="cuda:%d" % d) for d in devs]
    outs = [torch.empty_like(b) for b in bufs]
    with pytest.raises(RuntimeError):
        t.forwardSplitDevice([b.data_ptr() for b in bufs], [b.data_ptr() for b in outs], n, 10)
    ctx.close()


# ----------------------------------------------------------------------------------------------------------------
# 2-D FWT / WPT: BasicTransform.java:330-474 rows-then-columns / columns-then-rows
# ----------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("kind", ["fwt", "wpt"])
@pytest.mark.parametrize("cls,rows,cols,lvl_m,lvl_n,batch", [
    ("Haar1", 8, 8, 3, 3, 1),
    ("Daubechies4", 64, 128, 6, 7, 3),
    ("Daubechies4", 256, 64, 3, 2, 2),
    ("Daubechies8", 512, 1024, 9, 10, 2),
    ("Symlet8", 1024, 256, 4, 0, 1),
    ("Symlet8", 128, 2048, 0, 5, 2),
    ("Daubechies20", 16, 32, 4, 5, 2),      # filter (40 taps) longer than both dimensions
    ("Symlet10", 2, 4, 1, 2, 5),
    ("Daubechies2", 1, 64, 0, 6, 2),        # a single row: the column pass has nothing to do
    ("Daubechies3", 4096, 32, 12, 5, 1),
    ("Haar1", 512, 8, 9, 3, 3),             # fused column launches on a partial 32-column strip
    ("Daubechies2", 1024, 2, 10, 1, 2),
    ("Symlet10", 256, 96 // 3, 5, 2, 2),    # L = 20: two fused levels at a time
    ("Daubechies5", 128, 64, 7, 6, 2),      # smallest height the fused launches take
    ("Daubechies10", 2048, 64, 2, 6, 1),
    ("Daubechies4", 64, 1, 6, 0, 3),        # a single column
    ("Haar1", 2, 2, 1, 1, 1),
])
def test_2d_matches_oracle(jw, gpu_ctx, oracle, kind, cls, rows, cols, lvl_m, lvl_n, batch):
    w = jw.wavelets.create(cls)
    T = jw.CudaFastWaveletTransform if kind == "fwt" else jw.CudaWaveletPacketTransform
    t = T(w)
    X = splitmix_uniform(1234 + rows + cols, (batch, rows, cols))
    s, wv = w.getScalingDeComposition(), w.getWaveletDeComposition()
    ref = oracle.batch2d(kind, X, lvl_m, lvl_n, s, wv, nthreads=8)
    got = t.forward2DBatch(X, lvl_m, lvl_n)
    assert _maxerr(got, ref, X) <= TOL
    exact = t.forward2DBatch(X, lvl_m, lvl_n, flags=jw.FLAG_EXACT)
    assert np.array_equal(exact, ref)
    rref = oracle.batch2d(kind, ref, lvl_m, lvl_n, w.getScalingReConstruction(), w.getWaveletReConstruction(),
                          reverse=True, nthreads=8)
    back = t.reverse2DBatch(ref, lvl_m, lvl_n)
    assert _maxerr(back, rref, X) <= TOL
    assert np.array_equal(t.reverse2DBatch(ref, lvl_m, lvl_n, flags=jw.FLAG_EXACT), rref)
    assert _maxerr(t.reverse2DBatch(got, lvl_m, lvl_n), X, X) <= PR_TOL


# ----------------------------------------------------------------------------------------------------------------
# 3-D FWT / WPT: BasicTransform.java:487-640 -- the 2-D transform of every matrix, then the lines along the first axis
# ----------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("kind", ["fwt", "wpt"])
@pytest.mark.parametrize("cls,p,q,r,lvls,batch", [
    ("Haar1", 8, 8, 8, (3, 3, 3), 1),
    ("Daubechies4", 16, 32, 64, (5, 6, 4), 2),      # lvlP goes to the q axis, lvlQ to the r axis, lvlR to the p axis
    ("Daubechies4", 64, 64, 64, (6, 6, 6), 1),      # a cube at full depth: forward(double[][][])
    ("Symlet8", 128, 16, 32, (2, 3, 7), 2),         # deep first-axis transform: fused column launches on 512 columns
    ("Daubechies20", 4, 8, 16, (3, 4, 2), 3),       # filter longer than every dimension
    ("Daubechies2", 1, 16, 16, (4, 4, 0), 2),       # a single matrix: the first-axis pass has nothing to do
    ("Daubechies3", 256, 4, 8, (0, 0, 8), 1),       # only the first axis is transformed
    ("Daubechies6", 32, 1, 64, (0, 6, 5), 2),       # a single row per matrix
])
def test_3d_matches_oracle(jw, gpu_ctx, oracle, kind, cls, p, q, r, lvls, batch):
    w = jw.wavelets.create(cls)
    T = jw.CudaFastWaveletTransform if kind == "fwt" else jw.CudaWaveletPacketTransform
    t = T(w)
    X = splitmix_uniform(4321 + p + q + r, (batch, p, q, r))
    s, wv = w.getScalingDeComposition(), w.getWaveletDeComposition()
    ref = oracle.batch3d(kind, X, lvls[0], lvls[1], lvls[2], s, wv, nthreads=8)
    got = t.forward3DBatch(X, *lvls)
    assert _maxerr(got, ref, X) <= TOL
    assert np.array_equal(t.forward3DBatch(X, *lvls, flags=jw.FLAG_EXACT), ref)
    rref = oracle.batch3d(kind, ref, lvls[0], lvls[1], lvls[2], w.getScalingReConstruction(),
                          w.getWaveletReConstruction(), reverse=True, nthreads=8)
    assert _maxerr(t.reverse3DBatch(ref, *lvls), rref, X) <= TOL
    assert np.array_equal(t.reverse3DBatch(ref, *lvls, flags=jw.FLAG_EXACT), rref)
    assert _maxerr(t.reverse3DBatch(got, *lvls), X, X) <= PR_TOL


def test_3d_java_overloads_device_buffers_and_errors(jw, gpu_ctx, oracle):
    """forward(double[][][]) = the exponents of the three dimensions (BasicTransform.java:490-493); device-resident
    variant; errors as the 1-D calls (a non-cubic space at default depth fails like the reference: lvlP = log2(p) is
    applied to the q axis)."""
    import torch
    w = jw.wavelets.Daubechies4()
    t = jw.CudaFastWaveletTransform(w)
    X = splitmix_uniform(7, (16, 16, 16))
    s, wv = w.getScalingDeComposition(), w.getWaveletDeComposition()
    ref = oracle.batch3d("fwt", X[None], 4, 4, 4, s, wv)[0]
    assert _maxerr(t.forward(X), ref, X) <= TOL
    assert _maxerr(t.forward(X, 4, 4, 4), ref, X) <= TOL
    assert _maxerr(t.reverse(t.forward(X)), X, X) <= PR_TOL
    assert _maxerr(t.reverse(t.forward(X, 1, 2, 3), 1, 2, 3), X, X) <= PR_TOL
    with pytest.raises(jw.JWaveFailure, match="2\\^p"):
        t.forward(np.zeros((16, 12, 16)))
    with pytest.raises(jw.JWaveFailure, match="out of range"):
        t.forward(np.zeros((32, 8, 8)))          # lvlP = 5 > log2(q) = 3, as in the reference
    d_in = torch.from_numpy(np.ascontiguousarray(np.stack([X, -X]))).cuda()
    d_out = torch.empty_like(d_in)
    t.forward3DDevice(d_in.data_ptr(), d_out.data_ptr(), 2, 16, 16, 16, 4, 4, 4)
    torch.cuda.synchronize()
    assert _maxerr(d_out.cpu().numpy()[0], ref, X) <= TOL and _maxerr(d_out.cpu().numpy()[1], -ref, X) <= TOL
    d_back = torch.empty_like(d_in)
    t.reverse3DDevice(d_out.data_ptr(), d_back.data_ptr(), 2, 16, 16, 16, 4, 4, 4)
    torch.cuda.synchronize()
    assert _maxerr(d_back.cpu().numpy()[0], X, X) <= PR_TOL
    lib = jw._native.load()
    f = (ctypes.c_double * 2)(0.5, 0.5)
    buf = np.zeros(512)
    rc = lib.jwc_fwt3d_forward(gpu_ctx.handle, buf.ctypes.data, buf.ctypes.data, 1, 6, 8, 8, 1, 1, 1, f, f, 2, 0)
    assert rc == -1 and b"2^p" in lib.jwc_last_error()


def test_complex_overloads(jw, gpu_ctx, oracle):
    """BasicTransform.java:257-320 forward / reverse(Complex[]): N complex numbers = ONE real array of length 2N with
    real and imaginary parts interleaved, through the 1-D transform at full depth."""
    n = 256
    z = splitmix_uniform(3, (n,)) + 1j * splitmix_uniform(4, (n,))
    bulk = np.empty(2 * n)
    bulk[0::2], bulk[1::2] = z.real, z.imag
    for T, op in ((jw.CudaFastWaveletTransform, "fwt"), (jw.CudaWaveletPacketTransform, "wpt")):
        w = jw.wavelets.Daubechies4()
        t = T(w)
        ref = oracle.batch(op + "_fwd", bulk[None], 9, w.getScalingDeComposition(), w.getWaveletDeComposition())[0]
        got = t.forward(z)
        assert got.dtype == np.complex128 and got.shape == (n,)
        assert _maxerr(got.real, ref[0::2], bulk) <= TOL and _maxerr(got.imag, ref[1::2], bulk) <= TOL
        back = t.reverse(got)
        assert np.max(np.abs(back - z)) <= PR_TOL


def test_2d_java_overloads_and_errors(jw, gpu_ctx, oracle):
    """forward(double[][]) = full depth in both dimensions (BasicTransform.java:336-340); errors as the 1-D calls."""
    w = jw.wavelets.Daubechies4()
    t = jw.CudaFastWaveletTransform(w)
    X = splitmix_uniform(99, (32, 64))
    s, wv = w.getScalingDeComposition(), w.getWaveletDeComposition()
    ref = oracle.batch2d("fwt", X[None], 5, 6, s, wv)[0]
    assert _maxerr(t.forward(X), ref, X) <= TOL
    assert _maxerr(t.forward(X, 5, 6), ref, X) <= TOL
    assert _maxerr(t.reverse(t.forward(X)), X, X) <= PR_TOL
    assert _maxerr(t.reverse(t.forward(X, 2, 3), 2, 3), X, X) <= PR_TOL
    with pytest.raises(jw.JWaveFailure, match="2\\^p"):
        t.forward(np.zeros((24, 64)))
    with pytest.raises(jw.JWaveFailure, match="out of range"):
        t.forward(np.zeros((32, 64)), 6, 6)
    lib = jw._native.load()
    f = (ctypes.c_double * 2)(0.5, 0.5)
    buf = np.zeros(64)
    rc = lib.jwc_fwt2d_forward(gpu_ctx.handle, buf.ctypes.data, buf.ctypes.data, 1, 6, 8, 1, 1, f, f, 2, 0)
    assert rc == -1 and b"2^p" in lib.jwc_last_error()


def test_modwt_coefficients_format_from_device_layout(jw, gpu_ctx, oracle):
    """the device's row layout IS the MODWTCoefficients backing array and the flat forward() format."""
    w = jw.wavelets.Daubechies4()
    t = jw.CudaMODWTTransform(w)
    x = splitmix_uniform(31, (1024,))
    g, h = oracle.modwt_filters(w.getScalingDeComposition(), w.getWaveletDeComposition())
    ref = oracle.modwt_forward(x, 3, g, h)
    c = t.forwardMODWTCoefficients(x, 3)
    assert c.getTotalSize() == 4 * 1024
    for lvl in (1, 2, 3):
        assert _maxerr(c.getDetails(lvl), ref[lvl - 1], x) <= TOL
        assert _maxerr(c.getView(lvl).toArray(), ref[lvl - 1], x) <= TOL
    assert _maxerr(c.getApproximation(), ref[3], x) <= TOL
    assert np.array_equal(c.backingArray(), t.forward(x, 3))
    assert _maxerr(t.inverseMODWTCoefficients(c), x, x) <= PR_TOL
    assert _maxerr(t.reverse(c.backingArray(), 3), x, x) <= PR_TOL
    # the level-less reverse searches the SMALLEST 2^p N with total/N - 1 <= log2 N (MODWTTransform.java:888-897):
    # for 4 x 1024 coefficients that is N = 512, J = 7, not the shape that produced them -- same as the reference
    assert len(t.reverse(c.backingArray())) == 512


# ----------------------------------------------------------------------------------------------------------------
# arbitrary length: Ancient-Egyptian decomposition
# ----------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("kind", ["fwt", "wpt"])
@pytest.mark.parametrize("cls,n,batch", [
    ("Haar1", 42, 3),                 # 32 | 8 | 2
    ("Daubechies4", 127, 5),          # 64 | 32 | 16 | 8 | 4 | 2 | 1: odd row stride, every block misaligned
    ("Daubechies8", 1000, 4),         # 512 | 256 | 128 | 64 | 32 | 8
    ("Symlet8", 65536 + 4096 + 6, 3),
    ("Daubechies20", 3 * 4096, 2),    # 8192 | 4096, filter longer than nothing here but L = 40 kernels
    ("Coiflet2", 1, 4),
    ("Daubechies2", 1 << 12, 2),      # a power of two: one block, identical to the plain transform
    ("Daubechies3", 200000, 2),       # blocks above and below the short-signal tail threshold
    ("Daubechies4", 4097, 3),         # 4096 | 1 with an odd row stride: whole-signal / tail kernels without 16-byte alignment
    ("Haar1", 2048 + 1024 + 2, 4),
])
def test_ancient_egyptian_decomposition(jw, gpu_ctx, oracle, kind, cls, n, batch):
    w = jw.wavelets.create(cls)
    T = jw.CudaFastWaveletTransform if kind == "fwt" else jw.CudaWaveletPacketTransform
    aed = jw.AncientEgyptianDecomposition(T(w))
    X = splitmix_uniform(4242 + n, (batch, n))
    s, wv = w.getScalingDeComposition(), w.getWaveletDeComposition()
    ref = oracle.aed(kind, X, s, wv, nthreads=8)
    got = aed.forwardBatch(X)
    assert _maxerr(got, ref, X) <= TOL
    assert np.array_equal(aed.forwardBatch(X, flags=jw.FLAG_EXACT), ref)
    rref = oracle.aed(kind, ref, w.getScalingReConstruction(), w.getWaveletReConstruction(), reverse=True, nthreads=8)
    assert _maxerr(aed.reverseBatch(ref), rref, X) <= TOL
    assert np.array_equal(aed.reverseBatch(ref, flags=jw.FLAG_EXACT), rref)
    assert _maxerr(aed.reverse(aed.forward(X[0])), X[0], X) <= PR_TOL
    if n & (n - 1) == 0 and n > 1:
        assert np.array_equal(got, T(w).forwardBatch(X))


# ----------------------------------------------------------------------------------------------------------------
# sliding windows of one series (MODWTSlidingWindowTest.java:20-70)
# ----------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("cls,total,window,hop,J", [
    ("Haar1", 10000, 512, 64, 8),            # the reference test's configuration
    ("Daubechies4", 10000, 512, 64, 8),
    ("Daubechies4", 70000, 4096, 1000, 6),
    ("Symlet8", 5000, 300, 7, 4),            # window not a power of two, odd hop: element-wise loaders
    ("Daubechies2", 300000, 65536, 32768, 9),
    ("Daubechies20", 9000, 1024, 512, 5),
    ("Daubechies3", 64, 64, 5, 3),           # a single window
    ("Daubechies20", 40000, 3001, 777, 8),   # odd window longer than the whole-window kernels take, deep levels:
                                             # tile passes that walk the single cycle of the circular window
    ("Daubechies8", 50000, 6002, 1500, 9),   # window = 2 x odd: two cycles
])
def test_modwt_sliding_windows(jw, oracle, cls, total, window, hop, J):
    ctx = jw.Context([0])
    ctx.set_tuning("h2d_chunk_mb", 1)        # several pipeline chunks: overlapping input spans per chunk
    w = jw.wavelets.create(cls)
    t = jw.CudaMODWTTransform(w, context=ctx)
    x = chirp(1, total)[0] + 0.25 * splitmix_uniform(total + hop, (total,))
    nwin = (total - window) // hop + 1
    wins = np.lib.stride_tricks.sliding_window_view(x, window)[::hop][:nwin]
    ref, _ = _modwt_oracle(oracle, w, np.ascontiguousarray(wins), J)
    got = t.forwardMODWTWindows(x, window, hop, J)
    assert got.shape == (nwin, J + 1, window)
    assert _maxerr(got, ref, x) <= TOL
    assert np.array_equal(t.forwardMODWTWindows(x, window, hop, J, flags=jw.FLAG_EXACT), ref)
    assert _maxerr(t.inverseMODWTBatch(got), wins, x) <= PR_TOL
    ctx.close()


@pytest.mark.parametrize("shape,thr", [((1000,), 1.0), ((64, 513), 0.5), ((3, 7, 4096), 2.0), ((1,), 1.0),
                                       ((5_000_000,), 1.3)])
def test_compressor_magnitude(jw, gpu_ctx, oracle, shape, thr):
    """CompressorMagnitude.java:78-140 + Compressor.java:97-110: keep |c| >= mean|c| * threshold, zero the rest.
    Checked against the oracle's restatement (left-to-right sum, like the JVM): the device sum is a fixed-shape tree, so
    the magnitude agrees to O(1e-16) relative and the select can only differ for values within that of the cut."""
    x = splitmix_uniform(17 + len(shape), shape) * 3.0
    c = jw.CompressorMagnitude(thr)
    y = c.compress(x)
    exp, mag = oracle.compress_magnitude(x, thr)
    assert abs(c.getMagnitude() - mag) <= 1e-13 * mag
    cut = mag * thr
    differ = y != exp
    assert np.all(np.abs(np.abs(x[differ]) - cut) <= 1e-13 * max(cut, 1.0))
    assert np.count_nonzero(differ) <= 2
    sure = np.abs(np.abs(x) - cut) > 1e-12 * max(cut, 1.0)
    assert np.array_equal(y[sure], exp[sure])
    assert np.all((y == x) | (y == 0.0))
    assert c.calcCompressionRate(y) == pytest.approx(100.0 * np.count_nonzero(y == 0.0) / y.size)
    assert jw.CompressorMagnitude(-1.0).getThreshold() == 1.0


def test_windows_then_compressor_on_device(jw, gpu_ctx, oracle):
    """The reference's motivating chain on device buffers: sliding-window MODWT, then magnitude thresholding, one stream."""
    import torch
    w = jw.wavelets.Daubechies4()
    t = jw.CudaMODWTTransform(w)
    total, window, hop, J = 20000, 512, 64, 8
    x = chirp(1, total)[0]
    nwin = (total - window) // hop + 1
    dx = torch.from_numpy(x).cuda()
    dc = torch.empty((nwin, J + 1, window), dtype=torch.float64, device="cuda")
    dm = torch.zeros(1, dtype=torch.float64, device="cuda")
    lib = jw._native.load()
    g, h = t._filters()
    g = np.ascontiguousarray(g)
    h = np.ascontiguousarray(h)
    dp = ctypes.POINTER(ctypes.c_double)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream or 1)
    rc = lib.jwc_modwt_forward_windows_dev(gpu_ctx.handle, 0, st, ctypes.c_void_p(dx.data_ptr()),
                                           ctypes.c_void_p(dc.data_ptr()), total, window, hop, J,
                                           g.ctypes.data_as(dp), h.ctypes.data_as(dp), len(g), 0)
    assert rc == 0, lib.jwc_last_error()
    raw = dc.clone()
    rc = lib.jwc_compress_magnitude_dev(gpu_ctx.handle, 0, st, ctypes.c_void_p(dc.data_ptr()),
                                        ctypes.c_void_p(dc.data_ptr()), dc.numel(), 1.0, ctypes.c_void_p(dm.data_ptr()))
    assert rc == 0, lib.jwc_last_error()
    torch.cuda.synchronize()
    wins = np.lib.stride_tricks.sliding_window_view(x, window)[::hop][:nwin]
    ref, _ = _modwt_oracle(oracle, w, np.ascontiguousarray(wins), J)
    assert _maxerr(raw.cpu().numpy(), ref, x) <= TOL
    mag = float(np.mean(np.abs(ref)))
    assert abs(float(dm.item()) - mag) <= 1e-12 * mag
    y = dc.cpu().numpy()
    sure = np.abs(np.abs(ref) - mag) > 1e-9
    assert np.array_equal(y[sure] != 0.0, (np.abs(ref) >= mag)[sure])


@pytest.mark.parametrize("cls,window,hop,J,thr", [("Daubechies4", 512, 64, 8, 1.0), ("Haar1", 256, 32, 5, 0.5),
                                                  ("Symlet8", 1024, 100, 4, 2.0), ("Daubechies4", 4096, 512, 6, 1.0)])
def test_windows_compress_fused(jw, gpu_ctx, oracle, cls, window, hop, J, thr):
    """SURVEY 8f row 4 as worded: the thresholding fused into the transform's store.  jwc_modwt_forward_windows_compress_dev
    = window transform with sum |c| taken in the store epilogue + ONE select pass, against the oracle chain
    (forwardMODWT per window, then the CompressorMagnitude restatement over all coefficients).  The last shape is one the
    whole-window kernel declines (tile kernels + separate reduction): same result."""
    import torch
    w = jw.wavelets.create(cls)
    t = jw.CudaMODWTTransform(w)
    total = 40 * hop + window
    x = chirp(1, total)[0] + 0.1 * splitmix_uniform(5, (total,))
    nwin = (total - window) // hop + 1
    dx = torch.from_numpy(x).cuda()
    dc = torch.empty((nwin, J + 1, window), dtype=torch.float64, device="cuda")
    dm = torch.zeros(1, dtype=torch.float64, device="cuda")
    t.forwardMODWTWindowsCompressDevice(dx.data_ptr(), dc.data_ptr(), total, window, hop, J, thr, dm.data_ptr(),
                                        stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    wins = np.ascontiguousarray(np.lib.stride_tricks.sliding_window_view(x, window)[::hop][:nwin])
    ref, _ = _modwt_oracle(oracle, w, wins, J)
    exp, mag = oracle.compress_magnitude(ref, thr)
    assert abs(float(dm.item()) - mag) <= 1e-12 * mag
    y = dc.cpu().numpy()
    cut = mag * thr
    sure = np.abs(np.abs(ref) - cut) > 1e-9
    assert np.array_equal(y[sure] != 0.0, (exp != 0.0)[sure])
    kept = sure & (exp != 0.0)
    assert _maxerr(y[kept], exp[kept], x) <= TOL
    assert 0 < np.count_nonzero(y == 0.0) < y.size


@pytest.mark.parametrize("cls,n,lvl,batch,force", [
    ("Haar1", 1024, 2, 7, 0), ("Daubechies4", 2048, 3, 5, 0), ("Daubechies8", 4096, 4, 3, 0),
    ("Symlet10", 4096, 5, 2, 0), ("Daubechies4", 4096, 12, 3, 1), ("Daubechies6", 1024, 10, 4, 1),
])
def test_fwt_whole_signal_kernel(jw, oracle, cls, n, lvl, batch, force):
    """Signals of 1024..4096 samples: the whole signal in shared memory, levels in place (jwc_dwt_whole.cu); by default
    for shallow transforms, forced here for full depth as well."""
    ctx = jw.Context([0])
    ctx.set_tuning("dwt_whole", force)
    w = jw.wavelets.create(cls)
    t = jw.CudaFastWaveletTransform(w, context=ctx)
    X = splitmix_uniform(n + lvl, (batch, n))
    s, wv = w.getScalingDeComposition(), w.getWaveletDeComposition()
    ref = oracle.batch("fwt_fwd", X, lvl, s, wv, nthreads=8)
    l0 = ctx.launch_count()
    got = t.forwardBatch(X, lvl)
    assert ctx.launch_count() - l0 == 1          # one launch: the whole-signal kernel took it
    assert _maxerr(got, ref, X) <= TOL
    rref = oracle.batch("fwt_rev", ref, lvl, w.getScalingReConstruction(), w.getWaveletReConstruction(), nthreads=8)
    l0 = ctx.launch_count()
    back = t.reverseBatch(ref, lvl)
    assert ctx.launch_count() - l0 == 1
    assert _maxerr(back, rref, X) <= TOL
    assert _maxerr(t.reverseBatch(got, lvl), X, X) <= PR_TOL
    ctx.close()


@pytest.mark.parametrize("cls,n,lvl,batch,one_launch", [
    ("Symlet8", 4096, 6, 5, True), ("Haar1", 1024, 3, 9, True), ("Daubechies2", 64, 6, 33, True),
    ("Daubechies10", 2048, 5, 3, True), ("Daubechies4", 2048, 11, 2, False), ("Coiflet1", 256, 8, 4, True),
    ("Daubechies5", 4096, 4, 3, True),
])
def test_wpt_whole_signal_kernel(jw, oracle, cls, n, lvl, batch, one_launch):
    """Short packet transforms: every block of every level in place in shared memory (jwc_dwt_whole.cu, tree mode);
    depths whose busiest level has more work items than a CTA has threads fall back to the tile kernels."""
    ctx = jw.Context([0])
    ctx.set_tuning("dwt_whole", 1)     # also for the long filters, which default to the tile kernels
    w = jw.wavelets.create(cls)
    t = jw.CudaWaveletPacketTransform(w, context=ctx)
    X = splitmix_uniform(7 * n + lvl, (batch, n))
    s, wv = w.getScalingDeComposition(), w.getWaveletDeComposition()
    ref = oracle.batch("wpt_fwd", X, lvl, s, wv, nthreads=8)
    l0 = ctx.launch_count()
    got = t.forwardBatch(X, lvl)
    assert (ctx.launch_count() - l0 == 1) == one_launch
    assert _maxerr(got, ref, X) <= TOL
    rref = oracle.batch("wpt_rev", ref, lvl, w.getScalingReConstruction(), w.getWaveletReConstruction(), nthreads=8)
    back = t.reverseBatch(ref, lvl)
    assert _maxerr(back, rref, X) <= TOL
    assert _maxerr(t.reverseBatch(got, lvl), X, X) <= PR_TOL
    ctx.close()


@pytest.mark.parametrize("cls,n,lvl", [("Haar1", 1 << 16, 16), ("Daubechies8", 1 << 15, 9), ("Symlet5", 1 << 17, 4)])
def test_fwt_inverse_tiled_in_place_variant(jw, oracle, cls, n, lvl):
    """The opt-in tiled in-place inverse pass (dwt_tile_inv = 1; slower than the default TMA tile kernels, kept as an
    independent implementation): same results."""
    ctx = jw.Context([0])
    ctx.set_tuning("dwt_tile_inv", 1)
    w = jw.wavelets.create(cls)
    t = jw.CudaFastWaveletTransform(w, context=ctx)
    X = splitmix_uniform(n + lvl, (3, n))
    s, wv = w.getScalingDeComposition(), w.getWaveletDeComposition()
    ref = oracle.batch("fwt_fwd", X, lvl, s, wv, nthreads=8)
    rref = oracle.batch("fwt_rev", ref, lvl, w.getScalingReConstruction(), w.getWaveletReConstruction(), nthreads=8)
    assert _maxerr(t.reverseBatch(ref, lvl), rref, X) <= TOL
    assert np.array_equal(t.reverseBatch(ref, lvl), jw.CudaFastWaveletTransform(w).reverseBatch(ref, lvl)) or \
        _maxerr(t.reverseBatch(ref, lvl), jw.CudaFastWaveletTransform(w).reverseBatch(ref, lvl), X) <= TOL
    ctx.close()


@pytest.mark.parametrize("cls,n,J,batch", [("Daubechies20", 65536, 8, 3), ("Daubechies4", 65536, 11, 3),
                                            ("Symlet8", 1 << 17, 9, 2), ("Daubechies8", 40000, 6, 2)])
def test_modwt_cycle_walk_and_tile_wait_variants_agree(jw, oracle, cls, n, J, batch):
    """Two code paths that ordinary shapes do not reach on their own: (1) the cycle-walk instantiation of the fused MODWT
    passes (rows of a phase-split pass addressed as (i 2^j0) mod n -- what lengths 2^j0 does not divide run) forced onto
    lengths where it must give the plain phase addressing, (2) the round-1 form of the inverse tile wait (one thread on
    the mbarrier + a block barrier).  Both must reproduce the default kernels BIT FOR BIT: same arithmetic, only
    addresses / synchronisation differ."""
    w = jw.wavelets.create(cls)
    X = _inputs(n + J, batch, n)
    base = jw.CudaMODWTTransform(w)
    c0 = base.forwardMODWTBatch(X, J)
    x0 = base.inverseMODWTBatch(c0)
    ref, (g, h) = _modwt_oracle(oracle, w, X[:1], J)
    assert _maxerr(c0[:1], ref, X) <= TOL
    for key in ("modwt_force_wrap", "top_barrier"):
        ctx = jw.Context([0])
        ctx.set_tuning(key, 1)
        t = jw.CudaMODWTTransform(w, context=ctx)
        assert np.array_equal(t.forwardMODWTBatch(X, J), c0), key
        assert np.array_equal(t.inverseMODWTBatch(c0), x0), key
        ctx.close()


def test_fwt_wpt_tile_wait_variants_agree(jw):
    """Round-1 tile wait (top_barrier = 1) against the default in the FWT / WPT inverse tile kernels: bit-identical."""
    X = _inputs(5, 3, 1 << 17)
    for T, cls, lvl in ((jw.CudaFastWaveletTransform, "Daubechies8", 17), (jw.CudaFastWaveletTransform, "Haar1", 17),
                        (jw.CudaWaveletPacketTransform, "Symlet8", 6)):
        w = jw.wavelets.create(cls)
        c = T(w).forwardBatch(X, lvl)
        x0 = T(w).reverseBatch(c, lvl)
        ctx = jw.Context([0])
        ctx.set_tuning("top_barrier", 1)
        assert np.array_equal(T(w, context=ctx).reverseBatch(c, lvl), x0), (cls, lvl)
        ctx.close()


@pytest.mark.parametrize("cls,n,lvl,group", [("Daubechies8", 1 << 17, 17, 0), ("Haar1", 1 << 17, 17, 6),
                                              ("Daubechies20", 1 << 16, 9, 4), ("Coiflet5", 1 << 18, 7, 5),
                                              ("Symlet3", 1 << 15, 15, 8)])
def test_fwt_inverse_detail_tiles_requested_up_front(jw, oracle, cls, n, lvl, group):
    """The pyramid inverse requests every detail tile of a pass in its prologue (dwt_upfront, the default); with the
    one-level-ahead prefetch of round 1 (dwt_upfront = 0) the result is bit-identical, also for deeper passes than the
    planner's (dwt_group), and both match the oracle."""
    X = _inputs(11, 3, n)
    w = jw.wavelets.create(cls)
    c = jw.CudaFastWaveletTransform(w).forwardBatch(X, lvl)
    outs = []
    for up in (1, 0):
        ctx = jw.Context([0])
        ctx.set_tuning("dwt_upfront", up)
        if group:
            ctx.set_tuning("dwt_group", group)
        outs.append(jw.CudaFastWaveletTransform(w, context=ctx).reverseBatch(c, lvl))
        ctx.close()
    assert np.array_equal(outs[0], outs[1]), (cls, group)
    rref = oracle.batch("fwt_rev", c, lvl, w.getScalingReConstruction(), w.getWaveletReConstruction(), nthreads=8)
    assert _maxerr(outs[0], rref, X) <= TOL
    # (no perfect-reconstruction bound here: the reference's Coiflet5 table reconstructs to 3e-8 only, oracle included)
    assert _maxerr(outs[0], X, X) <= max(PR_TOL, 2.0 * _maxerr(rref, X, X))


def test_workspace_arenas_under_concurrent_device_calls(jw, oracle):
    """Multi-pass transforms take their workspace from the context's per-stream arenas.  Six host threads enqueue 2-D
    FWTs and deep 1-D FWTs on ONE context at once -- three on the context's shared stream, three on private streams --
    and every result must be right; then the arenas are released and the context still works."""
    import torch
    ctx = jw.Context([0])
    w = jw.wavelets.Daubechies4()
    t = jw.CudaFastWaveletTransform(w, context=ctx)
    s, wv = w.getScalingDeComposition(), w.getWaveletDeComposition()
    X2 = splitmix_uniform(3, (4, 256, 512))
    ref2 = oracle.batch2d("fwt", X2, 8, 9, s, wv, nthreads=8)
    X1 = splitmix_uniform(4, (6, 1 << 16))
    ref1 = oracle.batch("fwt_fwd", X1, 16, s, wv, nthreads=8)
    dX2, dX1 = torch.from_numpy(X2).cuda(), torch.from_numpy(X1).cuda()
    torch.cuda.synchronize()
    errs = []

    def work(tid):
        try:
            st = torch.cuda.Stream() if tid % 2 else None
            stream = st.cuda_stream if st is not None else None
            for it in range(8):
                o2, o1 = torch.empty_like(dX2), torch.empty_like(dX1)
                kw = {} if stream is None else {"stream": stream}
                if stream is None:   # the context's own stream: pass NULL through the raw ABI
                    lib = jw._native.load()
                    dp = ctypes.POINTER(ctypes.c_double)
                    f0, f1 = np.ascontiguousarray(s), np.ascontiguousarray(wv)
                    rc = lib.jwc_fwt2d_forward_dev(ctx.handle, 0, None, ctypes.c_void_p(dX2.data_ptr()),
                                                   ctypes.c_void_p(o2.data_ptr()), 4, 256, 512, 8, 9,
                                                   f0.ctypes.data_as(dp), f1.ctypes.data_as(dp), len(f0), 0)
                    rc |= lib.jwc_fwt_forward_dev(ctx.handle, 0, None, ctypes.c_void_p(dX1.data_ptr()),
                                                  ctypes.c_void_p(o1.data_ptr()), 6, 1 << 16, 16,
                                                  f0.ctypes.data_as(dp), f1.ctypes.data_as(dp), len(f0), 0)
                    assert rc == 0, lib.jwc_last_error()
                    ctx.synchronize()
                else:
                    t.forward2DDevice(dX2.data_ptr(), o2.data_ptr(), 4, 256, 512, 8, 9, **kw)
                    t.forwardDevice(dX1.data_ptr(), o1.data_ptr(), 6, 1 << 16, 16, **kw)
                    st.synchronize()
                if _maxerr(o2.cpu().numpy(), ref2, X2) > TOL or _maxerr(o1.cpu().numpy(), ref1, X1) > TOL:
                    errs.append((tid, it))
        except Exception as e:   # noqa: BLE001
            errs.append((tid, repr(e)))

    th = [threading.Thread(target=work, args=(i,)) for i in range(6)]
    [x_.start() for x_ in th]
    [x_.join() for x_ in th]
    assert not errs, errs[:3]
    ctx.release_scratch()
    assert _maxerr(t.forward2DBatch(X2, 8, 9), ref2, X2) <= TOL
    ctx.close()


def test_mixed_shapes_from_many_threads(jw, gpu_ctx, oracle):
    """Eight host threads on the default context, each with its own transform, wavelet and shape (different kernel
    instantiations, shared-memory sizes and pass plans), through the host-buffer API."""
    jobs = [("modwt", "Daubechies4", 4096, 6), ("modwt", "Daubechies20", 8192, 5), ("fwt", "Haar1", 1 << 15, 15),
            ("fwt", "Daubechies8", 1 << 14, 14), ("wpt", "Symlet8", 4096, 6), ("wpt", "Daubechies2", 1 << 13, 5),
            ("fwt", "Symlet10", 2048, 3), ("modwt", "Haar1", 512, 8)]
    refs, errs = [], []
    for kind, cls, n, lvl in jobs:
        w = jw.wavelets.create(cls)
        X = splitmix_uniform(n + lvl, (5, n))
        if kind == "modwt":
            ref, _ = _modwt_oracle(oracle, w, X, lvl)
        else:
            ref = oracle.batch(kind + "_fwd", X, lvl, w.getScalingDeComposition(), w.getWaveletDeComposition(), nthreads=4)
        refs.append((w, X, ref))

    def work(i):
        kind, cls, n, lvl = jobs[i]
        w, X, ref = refs[i]
        try:
            t = {"modwt": jw.CudaMODWTTransform, "fwt": jw.CudaFastWaveletTransform,
                 "wpt": jw.CudaWaveletPacketTransform}[kind](w)
            for it in range(12):
                got = t.forwardMODWTBatch(X, lvl) if kind == "modwt" else t.forwardBatch(X, lvl)
                back = t.inverseMODWTBatch(got) if kind == "modwt" else t.reverseBatch(got, lvl)
                if _maxerr(got, ref, X) > TOL or _maxerr(back, X, X) > PR_TOL:
                    errs.append((i, it))
        except Exception as e:   # noqa: BLE001
            errs.append((i, repr(e)))

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    [x_.start() for x_ in th]
    [x_.join() for x_ in th]
    assert not errs, errs[:3]


def test_device_pointers_that_are_only_8_byte_aligned(jw, gpu_ctx, oracle):
    """Views into larger device arrays start on any 8-byte boundary: every vectorised / bulk / cp.async16 path must
    notice and fall back to its element-wise loaders."""
    import torch
    w = jw.wavelets.Daubechies4()
    s, wv = w.getScalingDeComposition(), w.getWaveletDeComposition()
    f = jw.CudaFastWaveletTransform(w)
    m = jw.CudaMODWTTransform(w)

    def odd_view(arr):      # same values, data pointer = 16k + 8
        big = torch.empty(arr.size + 1, dtype=torch.float64, device="cuda")
        big[1:] = torch.from_numpy(arr.reshape(-1)).cuda()
        assert big[1:].data_ptr() % 16 == 8
        return big, big[1:]

    def odd_out(count):
        big = torch.empty(count + 1, dtype=torch.float64, device="cuda")
        return big, big[1:]

    # 2-D (fused column launches), short-signal FWT (whole-signal kernel), long FWT (bulk-copy tiles), MODWT tiles, windows
    X2 = splitmix_uniform(1, (2, 256, 64))
    _, dx = odd_view(X2); _, do = odd_out(X2.size)
    f.forward2DDevice(dx.data_ptr(), do.data_ptr(), 2, 256, 64, 8, 6)
    torch.cuda.synchronize()
    assert _maxerr(do.cpu().numpy().reshape(X2.shape), oracle.batch2d("fwt", X2, 8, 6, s, wv), X2) <= TOL
    for n, lvl in ((2048, 3), (1 << 16, 16), (300 * 0 + 512, 9)):
        X = splitmix_uniform(n, (3, n))
        _, dx = odd_view(X); _, do = odd_out(X.size)
        f.forwardDevice(dx.data_ptr(), do.data_ptr(), 3, n, lvl)
        torch.cuda.synchronize()
        ref = oracle.batch("fwt_fwd", X, lvl, s, wv)
        assert _maxerr(do.cpu().numpy().reshape(X.shape), ref, X) <= TOL
        _, dc = odd_view(ref); _, dr = odd_out(X.size)
        f.reverseDevice(dc.data_ptr(), dr.data_ptr(), 3, n, lvl)
        torch.cuda.synchronize()
        assert _maxerr(dr.cpu().numpy().reshape(X.shape), X, X) <= PR_TOL
    for n, J in ((8192, 6), (512, 8)):
        X = splitmix_uniform(n + 1, (3, n))
        ref, _ = _modwt_oracle(oracle, w, X, J)
        _, dx = odd_view(X); _, do = odd_out(ref.size)
        m.forwardMODWTDevice(dx.data_ptr(), do.data_ptr(), 3, n, J)
        torch.cuda.synchronize()
        assert _maxerr(do.cpu().numpy().reshape(ref.shape), ref, X) <= TOL
        _, dr = odd_out(X.size)
        m.inverseMODWTDevice(do.data_ptr(), dr.data_ptr(), 3, n, J)
        torch.cuda.synchronize()
        assert _maxerr(dr.cpu().numpy().reshape(X.shape), X, X) <= PR_TOL
