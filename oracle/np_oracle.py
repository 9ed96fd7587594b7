"""oracle/np_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Second, independent restatement (numpy, written from the reference's Java, not from
jwave_oracle.c) of the JWave-Pro filter-bank path.  Every multiply and add is a separate numpy
element-wise op in the reference's order, so results must be bit-identical to the C oracle;
tests/test_oracle_golden.py asserts that.  Parity status: pinned by the same reference KATs as the
C oracle (see the header of jwave_oracle.c).  Paths relative to
/root/reference/src/main/java/jwave/.
"""
import math

import numpy as np


def build_orthonormal(scaling):
    """transforms/wavelets/Wavelet.java:104-122"""
    s = np.asarray(scaling, dtype=np.float64)
    L = len(s)
    w = np.empty(L)
    for i in range(L):
        w[i] = s[L - 1 - i] if i % 2 == 0 else -s[L - 1 - i]
    return w


def modwt_filters(scaling, wavelet):
    """transforms/MODWTTransform.java:452-484 + normalize :599-606"""
    def normalize(f):
        f = np.array(f, dtype=np.float64)
        energy = 0.0
        for c in f:
            energy += float(c) * float(c)
        norm = math.sqrt(energy)
        if norm > 1e-12:
            f = f / norm
        return f
    sf = math.sqrt(2.0)
    return normalize(scaling) / sf, normalize(wavelet) / sf


def _conv(x, f, stride, adjoint):
    """transforms/MODWTTransform.java:677-716 with the structural zeros of the upsampled filter skipped."""
    N = len(x)
    n = np.arange(N, dtype=np.int64)
    acc = np.zeros(N)
    for m in range(len(f)):
        idx = (n + m * stride) % N if adjoint else (n - m * stride) % N
        acc = acc + x[idx] * f[m]
    return acc


def modwt_forward(x, J, g, h):
    """transforms/MODWTTransform.java:256-306 -> array (J+1, N): W_1..W_J, V_J"""
    v = np.array(x, dtype=np.float64)
    out = np.empty((J + 1, len(v)))
    for j in range(1, J + 1):
        st = 1 << (j - 1)
        out[j - 1] = _conv(v, h, st, False)
        v = _conv(v, g, st, False)
    out[J] = v
    return out


def modwt_inverse(coeffs, g, h):
    """transforms/MODWTTransform.java:337-375"""
    coeffs = np.asarray(coeffs, dtype=np.float64)
    J = coeffs.shape[0] - 1
    v = coeffs[J].copy()
    for j in range(J, 0, -1):
        st = 1 << (j - 1)
        v = _conv(v, g, st, True) + _conv(coeffs[j - 1], h, st, True)
    return v


def wavelet_forward(x, length, s, w):
    """transforms/wavelets/Wavelet.java:236-260"""
    h = length >> 1
    i = np.arange(h, dtype=np.int64)
    lo = np.zeros(h)
    hi = np.zeros(h)
    for j in range(len(s)):
        k = (2 * i + j) % length
        lo = lo + x[k] * s[j]
        hi = hi + x[k] * w[j]
    return np.concatenate([lo, hi])


def wavelet_reverse(c, length, sr, wr):
    """transforms/wavelets/Wavelet.java:277-303 (scatter-add, i outer / j inner; plain loops)."""
    out = [0.0] * length
    h = length >> 1
    c = [float(v) for v in c[:length]]
    sr = [float(v) for v in sr]
    wr = [float(v) for v in wr]
    L = len(sr)
    for i in range(h):
        for j in range(L):
            k = (2 * i + j) % length
            out[k] += (c[i] * sr[j]) + (c[i + h] * wr[j])
    return np.array(out)


def fwt_forward(x, level, s, w):
    """transforms/FastWaveletTransform.java:71-101"""
    a = np.array(x, dtype=np.float64)
    h, l = len(a), 0
    while h >= 2 and l < level:
        a[:h] = wavelet_forward(a, h, s, w)
        h >>= 1
        l += 1
    return a


def fwt_reverse(c, level, sr, wr):
    """transforms/FastWaveletTransform.java:119-153"""
    a = np.array(c, dtype=np.float64)
    N = len(a)
    steps = int(round(math.log2(N))) if N > 0 else 0
    h = 2
    for _ in range(level, steps):
        h <<= 1
    while 2 <= h <= N:
        a[:h] = wavelet_reverse(a, h, sr, wr)
        h <<= 1
    return a


def wpt_forward(x, level, s, w):
    """transforms/WaveletPacketTransform.java:73-124"""
    a = np.array(x, dtype=np.float64)
    N = len(a)
    h, l = N, 0
    while h >= 2 and l < level:
        for p in range(N // h):
            a[p * h:(p + 1) * h] = wavelet_forward(a[p * h:(p + 1) * h].copy(), h, s, w)
        h >>= 1
        l += 1
    return a


def wpt_reverse(c, level, sr, wr):
    """transforms/WaveletPacketTransform.java:141-191"""
    a = np.array(c, dtype=np.float64)
    N = len(a)
    steps = int(round(math.log2(N))) if N > 0 else 0
    h = 2
    for _ in range(level, steps):
        h <<= 1
    while 2 <= h <= N:
        for p in range(N // h):
            a[p * h:(p + 1) * h] = wavelet_reverse(a[p * h:(p + 1) * h].copy(), h, sr, wr)
        h <<= 1
    return a
