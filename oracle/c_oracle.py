"""oracle/c_oracle.py -- TEST INFRASTRUCTURE: ctypes loader for oracle/libjwave_oracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libjwave_oracle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    src = os.path.join(_HERE, "jwave_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B" if force else "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.jwo_batch.argtypes = [ctypes.c_int, _dp, _dp, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                   _dp, _dp, ctypes.c_int, ctypes.c_int]
        _lib.jwo_batch.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def build_orthonormal(scaling):
    s = _c(scaling)
    w = np.empty_like(s)
    lib().jwo_build_orthonormal(_p(s), ctypes.c_int(len(s)), _p(w))
    return w


def modwt_filters(scaling, wavelet):
    s, w = _c(scaling), _c(wavelet)
    g, h = np.empty_like(s), np.empty_like(s)
    lib().jwo_modwt_filters(_p(s), _p(w), ctypes.c_int(len(s)), _p(g), _p(h))
    return g, h


def upsample(f, level):
    f = _c(f)
    M = (len(f) - 1) * (1 << max(level - 1, 0)) + 1
    out = np.empty(M)
    n = lib().jwo_upsample(_p(f), ctypes.c_int(len(f)), ctypes.c_int(level), _p(out))
    return out[:n]


def circular_convolve(x, f, adjoint=False):
    x, f = _c(x), _c(f)
    out = np.empty_like(x)
    fn = lib().jwo_circular_convolve_adjoint if adjoint else lib().jwo_circular_convolve
    fn(_p(x), ctypes.c_int(len(x)), _p(f), ctypes.c_int(len(f)), _p(out))
    return out


def fft(re, im=None, inverse=False):
    """FastFourierTransform.java:130-163 (Cooley-Tukey for 2^p, Bluestein otherwise); returns (re, im)."""
    r = _c(re).copy()
    i = np.zeros_like(r) if im is None else _c(im).copy()
    lib().jwo_fft(_p(r), _p(i), ctypes.c_int(len(r)), ctypes.c_int(1 if inverse else 0))
    return r, i


def modwt_forward(x, J, g, h, dense=False, fft=False):
    x, g, h = _c(x), _c(g), _c(h)
    N = len(x)
    out = np.empty((J + 1, N))
    if fft:
        rc = lib().jwo_modwt_forward_fft(_p(x), ctypes.c_int(N), ctypes.c_int(J), _p(g), _p(h), ctypes.c_int(len(g)), _p(out))
    else:
        rc = lib().jwo_modwt_forward(_p(x), ctypes.c_int(N), ctypes.c_int(J), _p(g), _p(h), ctypes.c_int(len(g)), _p(out),
                                     ctypes.c_int(1 if dense else 0))
    assert rc == 0, rc
    return out


def modwt_inverse(coeffs, g, h, dense=False, fft=False):
    c, g, h = _c(coeffs), _c(g), _c(h)
    J, N = c.shape[0] - 1, c.shape[1]
    x = np.empty(N)
    if fft:
        rc = lib().jwo_modwt_inverse_fft(_p(c), ctypes.c_int(N), ctypes.c_int(J), _p(g), _p(h), ctypes.c_int(len(g)), _p(x))
    else:
        rc = lib().jwo_modwt_inverse(_p(c), ctypes.c_int(N), ctypes.c_int(J), _p(g), _p(h), ctypes.c_int(len(g)), _p(x),
                                     ctypes.c_int(1 if dense else 0))
    assert rc == 0, rc
    return x


def wavelet_forward(x, length, s, w):
    x, s, w = _c(x), _c(s), _c(w)
    out = np.empty(length)
    lib().jwo_wavelet_forward(_p(x), ctypes.c_int(length), _p(s), _p(w), ctypes.c_int(len(s)), _p(out))
    return out


def wavelet_reverse(c, length, sr, wr):
    c, sr, wr = _c(c), _c(sr), _c(wr)
    out = np.empty(length)
    lib().jwo_wavelet_reverse(_p(c), ctypes.c_int(length), _p(sr), _p(wr), ctypes.c_int(len(sr)), _p(out))
    return out


def _tree(fn, x, level, f0, f1):
    x, f0, f1 = _c(x), _c(f0), _c(f1)
    out = np.empty_like(x)
    rc = fn(_p(x), ctypes.c_int(len(x)), ctypes.c_int(level), _p(f0), _p(f1), ctypes.c_int(len(f0)), _p(out))
    assert rc == 0, rc
    return out


def fwt_forward(x, level, s, w):
    return _tree(lib().jwo_fwt_forward, x, level, s, w)


def fwt_reverse(c, level, sr, wr):
    return _tree(lib().jwo_fwt_reverse, c, level, sr, wr)


def wpt_forward(x, level, s, w):
    return _tree(lib().jwo_wpt_forward, x, level, s, w)


def wpt_reverse(c, level, sr, wr):
    return _tree(lib().jwo_wpt_reverse, c, level, sr, wr)


OPS = {"modwt_fwd": 0, "modwt_inv": 1, "modwt_fwd_fft": 2, "modwt_inv_fft": 3,
       "fwt_fwd": 4, "fwt_rev": 5, "wpt_fwd": 6, "wpt_rev": 7}


def batch(op, x, level, f0, f1, nthreads=1):
    """x: (batch, N) (or (batch, level+1, N) for MODWT inverse). Returns the batched result."""
    x, f0, f1 = _c(x), _c(f0), _c(f1)
    code = OPS[op]
    if code in (1, 3):
        B, N = x.shape[0], x.shape[2]
        out = np.empty((B, N))
    elif code in (0, 2):
        B, N = x.shape
        out = np.empty((B, level + 1, N))
    else:
        B, N = x.shape
        out = np.empty((B, N))
    rc = lib().jwo_batch(code, _p(x), _p(out), B, N, level, _p(f0), _p(f1), len(f0), nthreads)
    assert rc == 0, rc
    return out


def parallel_wpt(x, level, f0, f1, reverse=False, nthreads=1):
    """ParallelWaveletPacketTransform.forward / reverse (ParallelWaveletPacketTransform.java:79-147) on every row of x,
    one signal after the other, each level's packets forked over `nthreads` workers when packet >= 64 and packets >= 8
    (:155-158), leaves of at most 16 packets (:197-233).  f0 / f1: DeCom filters forward, ReCon filters reverse."""
    x, f0, f1 = _c(x), _c(f0), _c(f1)
    B, N = x.shape
    out = np.empty((B, N))
    fn = lib().jwo_parallel_wpt
    fn.argtypes = [_dp, _dp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _dp, _dp, ctypes.c_int, ctypes.c_int,
                   ctypes.c_int]
    fn.restype = ctypes.c_int
    rc = fn(_p(x), _p(out), B, N, level, _p(f0), _p(f1), len(f0), nthreads, 1 if reverse else 0)
    assert rc == 0, rc
    return out


def compress_magnitude(x, threshold=1.0):
    """CompressorMagnitude.compress (CompressorMagnitude.java:78-139 + Compressor.java:97-170) on an array of any rank:
    returns (compressed array, magnitude)."""
    x = _c(x)
    out = np.empty_like(x)
    fn = lib().jwo_compress_magnitude
    fn.argtypes = [_dp, ctypes.c_int64, ctypes.c_double, _dp]
    fn.restype = ctypes.c_double
    mag = fn(_p(x), x.size, float(threshold), _p(out))
    return out, float(mag)


def batch2d(kind, x, lvl_m, lvl_n, f0, f1, reverse=False, nthreads=1):
    """2-D FWT / WPT of every matrix of x (batch, rows, cols), composed from the 1-D oracle exactly as the reference
    composes it: transforms/BasicTransform.java:361-399 (forward: every row with lvl_n, then every column of the result
    with lvl_m) and :436-474 (reverse: every column with lvl_m, then every row with lvl_n).  The reference copies each
    column into a temporary array; a transpose is the same gather."""
    x = _c(x)
    B, rows, cols = x.shape
    op = kind + ("_rev" if reverse else "_fwd")

    def along_rows(a, lvl):
        return batch(op, a.reshape(B * rows, cols), lvl, f0, f1, nthreads).reshape(B, rows, cols)

    def along_cols(a, lvl):
        t = np.ascontiguousarray(a.transpose(0, 2, 1)).reshape(B * cols, rows)
        r = batch(op, t, lvl, f0, f1, nthreads).reshape(B, cols, rows)
        return np.ascontiguousarray(r.transpose(0, 2, 1))

    if not reverse:
        return along_cols(along_rows(x, lvl_n), lvl_m)
    return along_rows(along_cols(x, lvl_m), lvl_n)


def batch3d(kind, x, lvl_p, lvl_q, lvl_r, f0, f1, reverse=False, nthreads=1):
    """3-D FWT / WPT of every space of x (batch, p, q, r), composed as the reference composes it:
    transforms/BasicTransform.java:509-565 forward(double[][][], lvlP, lvlQ, lvlR) -- the 2-D forward(mat, lvlP, lvlQ) of
    every matrix x[b][i] (so the rows of length r get lvlQ, the columns of length q get lvlP), then every line along the
    first axis with lvlR -- and :602-640 reverse, which keeps that order (2-D reverse first, then the first axis)."""
    x = _c(x)
    B, P, Q, R = x.shape
    op = kind + ("_rev" if reverse else "_fwd")
    mats = batch2d(kind, x.reshape(B * P, Q, R), lvl_p, lvl_q, f0, f1, reverse=reverse, nthreads=nthreads)
    lines = np.ascontiguousarray(mats.reshape(B, P, Q * R).transpose(0, 2, 1)).reshape(B * Q * R, P)
    out = batch(op, lines, lvl_r, f0, f1, nthreads).reshape(B, Q * R, P)
    return np.ascontiguousarray(out.transpose(0, 2, 1)).reshape(B, P, Q, R)


def aed_blocks(n):
    """tools/MathToolKit.java:57-84 decompose(): exponents of the descending powers of two that sum to n (42 -> 5, 3, 1)."""
    assert n >= 1
    out = []
    cur = n
    while cur >= 1:
        p = cur.bit_length() - 1
        out.append(p)
        cur -= 1 << p
    return out


def aed(kind, x, f0, f1, reverse=False, nthreads=1):
    """transforms/AncientEgyptianDecomposition.java:97-181: every 2^p block of every signal of x (batch, n), n arbitrary,
    through the wrapped transform's forward(double[]) / reverse(double[]) = full depth p; results at the same positions."""
    x = _c(x)
    out = np.empty_like(x)
    off = 0
    for p in aed_blocks(x.shape[1]):
        ln = 1 << p
        out[:, off:off + ln] = batch(kind + ("_rev" if reverse else "_fwd"), x[:, off:off + ln], p, f0, f1, nthreads)
        off += ln
    return out
