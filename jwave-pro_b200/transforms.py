"""Host-side mirror of the reference's transform classes for the three GPU transforms.

The reference is Java and this image has no JDK, so the drop-in classes exist twice: as Java sources that bind
the C ABI through Panama FFM (java/jwave/transforms/cuda/*.java, not compilable here) and as this Python mirror
with the same class / method names, argument meaning, validation order and error messages, so that the tests
in tests/ read like the reference's own JUnit tests.  Both sit on the same C ABI (include/jwavecuda.h); neither
has a CPU path.

Reference (relative to /root/reference/src/main/java/jwave/transforms/):
  BasicTransform.java:99-157,671-697        abstract 1-D API, isBinary, calcExponent
  BasicTransform.java:330-474               2-D forward / reverse (rows then columns; columns then rows)
  WaveletTransform.java:77-182              full-depth defaults, decompose / recompose
  FastWaveletTransform.java:71-153          CudaFastWaveletTransform
  WaveletPacketTransform.java:73-191        CudaWaveletPacketTransform
  MODWTTransform.java:256-443,452-606,854-912   CudaMODWTTransform
"""
import ctypes
import math
import threading

import numpy as np

from . import _native
from .exceptions import IllegalArgumentException, JWaveFailure

_dp = ctypes.POINTER(ctypes.c_double)


def _as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return a.ctypes.data_as(_dp)


def _out_buffer(out, shape):
    """Result buffer of a batched call: a fresh array, or the caller's `out` after checking that native code may write
    `shape` float64 values into it (a float32, strided, read-only or undersized array would be silent heap corruption)."""
    shape = tuple(int(v) for v in shape)
    if out is None:
        return np.empty(shape, dtype=np.float64)
    if not isinstance(out, np.ndarray) or out.dtype != np.float64:
        raise IllegalArgumentException("out must be a numpy float64 array")
    if not out.flags["C_CONTIGUOUS"] or not out.flags["WRITEABLE"]:
        raise IllegalArgumentException("out must be C-contiguous and writeable")
    if tuple(out.shape) != shape:
        raise IllegalArgumentException("out has shape %s, expected %s" % (tuple(out.shape), shape))
    return out


class BasicTransform:
    """transforms/BasicTransform.java:42 -- only the 1-D surface the hot path touches."""

    _name = None

    def getName(self):
        return self._name

    @staticmethod
    def isBinary(number):
        # tools/MathToolKit.java:185
        return number > 0 and (number & (number - 1)) == 0

    def calcExponent(self, number):
        # BasicTransform.java:687-697 + MathToolKit.getExponent :202 ((int)(ln f / ln 2), exact for 2^0..2^30)
        if not self.isBinary(number):
            raise JWaveFailure("BasicTransform#calcExponent - given number is not binary: "
                               "2^p | pEN .. = 1, 2, 4, 8, 16, 32, .. ")
        return int(number).bit_length() - 1


class WaveletTransform(BasicTransform):
    """transforms/WaveletTransform.java:42"""

    def __init__(self, wavelet, context=None):
        self._wavelet = wavelet
        self._ctx = context

    def getWavelet(self):
        return self._wavelet

    def _context(self):
        if self._ctx is None:
            self._ctx = _native.default_context()
        return self._ctx

    def _call(self, fn_name, src, dst, batch, n, levels, f0, f1, flags):
        lib = _native.load()
        f0, f1 = _as_f64(f0), _as_f64(f1)
        rc = getattr(lib, fn_name)(self._context().handle, src.ctypes.data, dst.ctypes.data, batch, n, levels,
                                   _ptr(f0), _ptr(f1), len(f0), flags)
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (fn_name, rc, _native.last_error()))

    def _call_dev(self, fn_name, d_src, d_dst, batch, n, levels, f0, f1, flags, stream=0, slot=0):
        """Device-resident variant: d_src / d_dst are raw device addresses (e.g. torch.Tensor.data_ptr());
        stream is a cudaStream_t handle.  0 means the legacy default stream (what torch's default stream is), passed
        as cudaStreamLegacy (0x1) because a NULL stream asks the library for the context's own stream."""
        lib = _native.load()
        f0, f1 = _as_f64(f0), _as_f64(f1)
        rc = getattr(lib, fn_name + "_dev")(self._context().handle, slot, ctypes.c_void_p(stream if stream else 1),
                                            ctypes.c_void_p(d_src), ctypes.c_void_p(d_dst), batch, n, levels,
                                            _ptr(f0), _ptr(f1), len(f0), flags)
        if rc != 0:
            raise RuntimeError("%s_dev failed (%d): %s" % (fn_name, rc, _native.last_error()))

    # ---- full-depth defaults (WaveletTransform.java:77-112) --------------------------------------------------
    def forward(self, arrTime, level=None, lvlN=None, lvlR=None):
        """1-D: forward(arrTime[, level]).  A 2-D array selects the reference's matrix overloads
        forward(double[][]) / forward(double[][], lvlM, lvlN) (BasicTransform.java:330-399), a 3-D array its space
        overloads forward(double[][][]) / forward(double[][][], lvlP, lvlQ, lvlR) (:487-565)."""
        if np.iscomplexobj(arrTime) and np.ndim(arrTime) == 1:
            # BasicTransform.java:257-280 forward(Complex[]): the N complex numbers as ONE real array of length 2N,
            # real and imaginary parts interleaved, through forward(double[]) at full depth, and back
            bulk = np.ascontiguousarray(arrTime, dtype=np.complex128).view(np.float64)
            return np.ascontiguousarray(self.forward(bulk)).view(np.complex128)
        if np.ndim(arrTime) == 3:
            return self.forward3D(arrTime, level, lvlN, lvlR)
        if np.ndim(arrTime) == 2:
            return self.forward2D(arrTime, level, lvlN)
        if level is None:
            if not self.isBinary(len(arrTime)):
                raise JWaveFailure("WaveletTransform#forward - given array length is not 2^p | p E N ... = "
                                   "1, 2, 4, 8, 16, 32, .. please use the Ancient Egyptian Decomposition for any "
                                   "other array length!")
            level = self.calcExponent(len(arrTime))
        return self._forward_level(arrTime, level)

    def reverse(self, arrHilb, level=None, lvlN=None, lvlR=None):
        if np.iscomplexobj(arrHilb) and np.ndim(arrHilb) == 1:      # BasicTransform.java:297-320 reverse(Complex[])
            bulk = np.ascontiguousarray(arrHilb, dtype=np.complex128).view(np.float64)
            return np.ascontiguousarray(self.reverse(bulk)).view(np.complex128)
        if np.ndim(arrHilb) == 3:
            return self.reverse3D(arrHilb, level, lvlN, lvlR)
        if np.ndim(arrHilb) == 2:
            return self.reverse2D(arrHilb, level, lvlN)
        if level is None:
            if not self.isBinary(len(arrHilb)):
                raise JWaveFailure("WaveletTransform#reverse - given array length is not 2^p | p E N ... = "
                                   "1, 2, 4, 8, 16, 32, .. please use the Ancient Egyptian Decomposition for any "
                                   "other array length!")
            level = self.calcExponent(len(arrHilb))
        return self._reverse_level(arrHilb, level)

    def decompose(self, arrTime):
        # WaveletTransform.java:136-147
        length = len(arrTime)
        levels = self.calcExponent(length)
        return np.stack([self.forward(arrTime, p) for p in range(levels + 1)])

    def recompose(self, matDeComp, level):
        # WaveletTransform.java:173-182
        if level < 0 or level >= len(matDeComp):
            raise JWaveFailure("WaveletTransform#recompose - given level is out of range")
        return self.reverse(matDeComp[level], level)


class _CudaPyramidBase(WaveletTransform):
    """Shared by FWT and WPT: same validation (FastWaveletTransform.java:74-83, WaveletPacketTransform.java:76-84)."""

    _fn = None      # "jwc_fwt" / "jwc_wpt"
    _cls = None     # class name used in the reference's messages

    def _check(self, length, level, direction):
        if not self.isBinary(length):
            raise JWaveFailure("%s#%s - given array length is not 2^p | p E N ... = 1, 2, 4, 8, 16, 32, .. "
                               "please use the Ancient Egyptian Decomposition for any other array length!"
                               % (self._cls, direction))
        if level < 0 or level > self.calcExponent(length):
            raise JWaveFailure("%s#%s - given level is out of range for given array" % (self._cls, direction))

    def _forward_level(self, arrTime, level, flags=0):
        x = _as_f64(arrTime)
        self._check(len(x), level, "forward")
        out = np.empty_like(x)
        self._call(self._fn + "_forward", x, out, 1, len(x), level, self._wavelet.getScalingDeComposition(),
                   self._wavelet.getWaveletDeComposition(), flags)
        return out

    def _reverse_level(self, arrHilb, level, flags=0):
        c = _as_f64(arrHilb)
        self._check(len(c), level, "reverse")
        out = np.empty_like(c)
        self._call(self._fn + "_inverse", c, out, 1, len(c), level, self._wavelet.getScalingReConstruction(),
                   self._wavelet.getWaveletReConstruction(), flags)
        return out

    # ---- batched entry points (rows = independent signals), host buffers -------------------------------------
    def forwardBatch(self, matTime, level=None, flags=0, out=None):
        X = _as_f64(matTime)
        B, N = X.shape
        if level is None:
            level = self.calcExponent(N)
        self._check(N, level, "forward")
        out = _out_buffer(out, X.shape)
        self._call(self._fn + "_forward", X, out, B, N, level, self._wavelet.getScalingDeComposition(),
                   self._wavelet.getWaveletDeComposition(), flags)
        return out

    def reverseBatch(self, matHilb, level=None, flags=0, out=None):
        C = _as_f64(matHilb)
        B, N = C.shape
        if level is None:
            level = self.calcExponent(N)
        self._check(N, level, "reverse")
        out = _out_buffer(out, C.shape)
        self._call(self._fn + "_inverse", C, out, B, N, level, self._wavelet.getScalingReConstruction(),
                   self._wavelet.getWaveletReConstruction(), flags)
        return out

    # ---- 2-D (BasicTransform.java:330-474; ParallelTransform.java:70-91,222-271 is the same arithmetic) ------
    def _levels2d(self, rows, cols, lvlM, lvlN, direction):
        if lvlM is None:
            lvlM = self.calcExponent(rows)    # BasicTransform.java:336-338 / :412-414
        if lvlN is None:
            lvlN = self.calcExponent(cols)
        self._check(cols, lvlN, direction)    # the reference fails in the first row transform ...
        self._check(rows, lvlM, direction)    # ... or in the first column transform
        return lvlM, lvlN

    def _call2d(self, fn_name, src, dst, batch, rows, cols, lvlM, lvlN, f0, f1, flags):
        lib = _native.load()
        f0, f1 = _as_f64(f0), _as_f64(f1)
        rc = getattr(lib, fn_name)(self._context().handle, src.ctypes.data, dst.ctypes.data, batch, rows, cols, lvlM,
                                   lvlN, _ptr(f0), _ptr(f1), len(f0), flags)
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (fn_name, rc, _native.last_error()))

    def forward2DBatch(self, cubeTime, lvlM=None, lvlN=None, flags=0, out=None):
        """[batch][rows][cols] -> same shape; every matrix as forward(double[][], lvlM, lvlN)."""
        X = _as_f64(cubeTime)
        B, rows, cols = X.shape
        lvlM, lvlN = self._levels2d(rows, cols, lvlM, lvlN, "forward")
        out = _out_buffer(out, X.shape)
        self._call2d(self._fn + "2d_forward", X, out, B, rows, cols, lvlM, lvlN,
                     self._wavelet.getScalingDeComposition(), self._wavelet.getWaveletDeComposition(), flags)
        return out

    def reverse2DBatch(self, cubeHilb, lvlM=None, lvlN=None, flags=0, out=None):
        C = _as_f64(cubeHilb)
        B, rows, cols = C.shape
        lvlM, lvlN = self._levels2d(rows, cols, lvlM, lvlN, "reverse")
        out = _out_buffer(out, C.shape)
        self._call2d(self._fn + "2d_inverse", C, out, B, rows, cols, lvlM, lvlN,
                     self._wavelet.getScalingReConstruction(), self._wavelet.getWaveletReConstruction(), flags)
        return out

    def forward2D(self, matTime, lvlM=None, lvlN=None, flags=0):
        X = _as_f64(matTime)
        return self.forward2DBatch(X[None, :, :], lvlM, lvlN, flags)[0]

    def reverse2D(self, matHilb, lvlM=None, lvlN=None, flags=0):
        C = _as_f64(matHilb)
        return self.reverse2DBatch(C[None, :, :], lvlM, lvlN, flags)[0]

    # ---- 3-D (BasicTransform.java:487-640) --------------------------------------------------------------------
    def _levels3d(self, p, q, r, lvlP, lvlQ, lvlR, direction):
        if lvlP is None:      # BasicTransform.java:490-493 / :582-585: the exponents of the three dimensions, in this order
            lvlP = self.calcExponent(p)
        if lvlQ is None:
            lvlQ = self.calcExponent(q)
        if lvlR is None:
            lvlR = self.calcExponent(r)
        # the reference hands (lvlP, lvlQ) to the 2-D transform of every [q][r] matrix -- rows of length r with lvlQ,
        # columns of length q with lvlP -- and lvlR to the lines of length p along the first axis (:532, :555)
        self._check(r, lvlQ, direction)
        self._check(q, lvlP, direction)
        self._check(p, lvlR, direction)
        return lvlP, lvlQ, lvlR

    def _call3d(self, fn_name, src, dst, batch, p, q, r, lvls, f0, f1, flags):
        lib = _native.load()
        f0, f1 = _as_f64(f0), _as_f64(f1)
        rc = getattr(lib, fn_name)(self._context().handle, src.ctypes.data, dst.ctypes.data, batch, p, q, r, lvls[0],
                                   lvls[1], lvls[2], _ptr(f0), _ptr(f1), len(f0), flags)
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (fn_name, rc, _native.last_error()))

    def forward3DBatch(self, spcTime, lvlP=None, lvlQ=None, lvlR=None, flags=0, out=None):
        """[batch][p][q][r] -> same shape; every space as forward(double[][][], lvlP, lvlQ, lvlR)."""
        X = _as_f64(spcTime)
        B, p, q, r = X.shape
        lvls = self._levels3d(p, q, r, lvlP, lvlQ, lvlR, "forward")
        out = _out_buffer(out, X.shape)
        self._call3d(self._fn + "3d_forward", X, out, B, p, q, r, lvls, self._wavelet.getScalingDeComposition(),
                     self._wavelet.getWaveletDeComposition(), flags)
        return out

    def reverse3DBatch(self, spcHilb, lvlP=None, lvlQ=None, lvlR=None, flags=0, out=None):
        C = _as_f64(spcHilb)
        B, p, q, r = C.shape
        lvls = self._levels3d(p, q, r, lvlP, lvlQ, lvlR, "reverse")
        out = _out_buffer(out, C.shape)
        self._call3d(self._fn + "3d_inverse", C, out, B, p, q, r, lvls, self._wavelet.getScalingReConstruction(),
                     self._wavelet.getWaveletReConstruction(), flags)
        return out

    def forward3D(self, spcTime, lvlP=None, lvlQ=None, lvlR=None, flags=0):
        X = _as_f64(spcTime)
        return self.forward3DBatch(X[None], lvlP, lvlQ, lvlR, flags)[0]

    def reverse3D(self, spcHilb, lvlP=None, lvlQ=None, lvlR=None, flags=0):
        C = _as_f64(spcHilb)
        return self.reverse3DBatch(C[None], lvlP, lvlQ, lvlR, flags)[0]

    def _call3d_dev(self, fn_name, d_src, d_dst, batch, p, q, r, lvls, f0, f1, flags, stream, slot):
        lib = _native.load()
        f0, f1 = _as_f64(f0), _as_f64(f1)
        rc = getattr(lib, fn_name + "_dev")(self._context().handle, slot, ctypes.c_void_p(stream if stream else 1),
                                            ctypes.c_void_p(d_src), ctypes.c_void_p(d_dst), batch, p, q, r, lvls[0],
                                            lvls[1], lvls[2], _ptr(f0), _ptr(f1), len(f0), flags)
        if rc != 0:
            raise RuntimeError("%s_dev failed (%d): %s" % (fn_name, rc, _native.last_error()))

    def forward3DDevice(self, d_in, d_out, batch, p, q, r, lvlP, lvlQ, lvlR, stream=0, flags=0, slot=0):
        lvls = self._levels3d(p, q, r, lvlP, lvlQ, lvlR, "forward")
        self._call3d_dev(self._fn + "3d_forward", d_in, d_out, batch, p, q, r, lvls,
                         self._wavelet.getScalingDeComposition(), self._wavelet.getWaveletDeComposition(), flags,
                         stream, slot)

    def reverse3DDevice(self, d_in, d_out, batch, p, q, r, lvlP, lvlQ, lvlR, stream=0, flags=0, slot=0):
        lvls = self._levels3d(p, q, r, lvlP, lvlQ, lvlR, "reverse")
        self._call3d_dev(self._fn + "3d_inverse", d_in, d_out, batch, p, q, r, lvls,
                         self._wavelet.getScalingReConstruction(), self._wavelet.getWaveletReConstruction(), flags,
                         stream, slot)

    def _call2d_dev(self, fn_name, d_src, d_dst, batch, rows, cols, lvlM, lvlN, f0, f1, flags, stream, slot):
        lib = _native.load()
        f0, f1 = _as_f64(f0), _as_f64(f1)
        rc = getattr(lib, fn_name + "_dev")(self._context().handle, slot, ctypes.c_void_p(stream if stream else 1),
                                            ctypes.c_void_p(d_src), ctypes.c_void_p(d_dst), batch, rows, cols, lvlM,
                                            lvlN, _ptr(f0), _ptr(f1), len(f0), flags)
        if rc != 0:
            raise RuntimeError("%s_dev failed (%d): %s" % (fn_name, rc, _native.last_error()))

    def forward2DDevice(self, d_in, d_out, batch, rows, cols, lvlM, lvlN, stream=0, flags=0, slot=0):
        lvlM, lvlN = self._levels2d(rows, cols, lvlM, lvlN, "forward")
        self._call2d_dev(self._fn + "2d_forward", d_in, d_out, batch, rows, cols, lvlM, lvlN,
                         self._wavelet.getScalingDeComposition(), self._wavelet.getWaveletDeComposition(), flags,
                         stream, slot)

    def reverse2DDevice(self, d_in, d_out, batch, rows, cols, lvlM, lvlN, stream=0, flags=0, slot=0):
        lvlM, lvlN = self._levels2d(rows, cols, lvlM, lvlN, "reverse")
        self._call2d_dev(self._fn + "2d_inverse", d_in, d_out, batch, rows, cols, lvlM, lvlN,
                         self._wavelet.getScalingReConstruction(), self._wavelet.getWaveletReConstruction(), flags,
                         stream, slot)

    # ---- device-resident (benchmarks, pipelines that keep data in HBM) ------------------------------------
    def forwardDevice(self, d_in, d_out, batch, n, level, stream=0, flags=0, slot=0):
        self._check(n, level, "forward")
        self._call_dev(self._fn + "_forward", d_in, d_out, batch, n, level, self._wavelet.getScalingDeComposition(),
                       self._wavelet.getWaveletDeComposition(), flags, stream, slot)

    def reverseDevice(self, d_in, d_out, batch, n, level, stream=0, flags=0, slot=0):
        self._check(n, level, "reverse")
        self._call_dev(self._fn + "_inverse", d_in, d_out, batch, n, level, self._wavelet.getScalingReConstruction(),
                       self._wavelet.getWaveletReConstruction(), flags, stream, slot)


    # ---- one long series split over the context's devices (halo exchange between ring neighbours) ------------------
    def _split(self, direction, src_ptrs, dst_ptrs, n, level):
        lib = _native.load()
        self._check(n, level, direction)
        P = self._context().num_devices()
        if len(src_ptrs) != P or len(dst_ptrs) != P:
            raise ValueError("need one chunk pointer per device of the context (%d)" % P)
        w = self._wavelet
        f0, f1 = ((w.getScalingDeComposition(), w.getWaveletDeComposition()) if direction == "forward"
                  else (w.getScalingReConstruction(), w.getWaveletReConstruction()))
        f0, f1 = _as_f64(f0), _as_f64(f1)
        a = (ctypes.c_void_p * P)(*src_ptrs)
        b = (ctypes.c_void_p * P)(*dst_ptrs)
        fn = getattr(lib, "%s_%s_split_dev" % (self._fn, "forward" if direction == "forward" else "inverse"))
        rc = fn(self._context().handle, a, b, n, level, _ptr(f0), _ptr(f1), len(f0), 0)
        if rc != 0:
            raise RuntimeError("split %s failed (%d): %s" % (direction, rc, _native.last_error()))

    def forwardSplitDevice(self, d_in_chunks, d_out_chunks, n, level):
        """d_in_chunks[p]: device address of x[n*p/P .. n*(p+1)/P) on device slot p; d_out_chunks[p]: n/P doubles in the
        local layout described in include/jwavecuda.h (splitLayoutToGlobal maps it back to the reference's array)."""
        self._split("forward", d_in_chunks, d_out_chunks, n, level)

    def reverseSplitDevice(self, d_in_chunks, d_out_chunks, n, level):
        self._split("reverse", d_in_chunks, d_out_chunks, n, level)

    def splitLevels(self, n, level):
        """levels of an n-sample transform that run split over this context's devices (the rest is the FWT remainder)"""
        return int(_native.load().jwc_dwt_split_levels(self._context().handle, n, level))

    def splitLayoutToGlobal(self, chunks, n, level):
        """Assemble the reference's coefficient array from the per-device chunks of forwardSplitDevice (numpy arrays)."""
        P = len(chunks)
        ln = n // P
        out = np.empty(n)
        if self._fn == "jwc_wpt":
            steps = min(level, self.calcExponent(n))
            part = ln >> steps
            for c in range(1 << steps):
                for p in range(P):
                    out[c * (n >> steps) + p * part:c * (n >> steps) + (p + 1) * part] = chunks[p][c * part:(c + 1) * part]
            return out
        ls = self.splitLevels(n, level)
        tp = ln >> ls
        for p in range(P):
            out[p * tp:(p + 1) * tp] = chunks[p][:tp]
        for l in range(ls, 0, -1):
            part = ln >> l
            for p in range(P):
                out[(n >> l) + p * part:(n >> l) + (p + 1) * part] = chunks[p][part:2 * part]
        return out

    def globalToSplitLayout(self, coeffs, P, level):
        """Inverse of splitLayoutToGlobal: the per-device chunks (list of numpy arrays) of a coefficient array."""
        coeffs = _as_f64(coeffs)
        n = len(coeffs)
        ln = n // P
        chunks = [np.empty(ln) for _ in range(P)]
        if self._fn == "jwc_wpt":
            steps = min(level, self.calcExponent(n))
            part = ln >> steps
            for c in range(1 << steps):
                for p in range(P):
                    chunks[p][c * part:(c + 1) * part] = coeffs[c * (n >> steps) + p * part:c * (n >> steps) + (p + 1) * part]
            return chunks
        ls = self.splitLevels(n, level)
        tp = ln >> ls
        for p in range(P):
            chunks[p][:tp] = coeffs[p * tp:(p + 1) * tp]
        for l in range(ls, 0, -1):
            part = ln >> l
            for p in range(P):
                chunks[p][part:2 * part] = coeffs[(n >> l) + p * part:(n >> l) + (p + 1) * part]
        return chunks


class CudaFastWaveletTransform(_CudaPyramidBase):
    """Drop-in for transforms/FastWaveletTransform.java (same _name, :52)."""

    _fn = "jwc_fwt"
    _cls = "FastWaveletTransform"

    def __init__(self, wavelet, context=None):
        super().__init__(wavelet, context)
        self._name = "Fast Wavelet Transform"


class CudaWaveletPacketTransform(_CudaPyramidBase):
    """Drop-in for transforms/WaveletPacketTransform.java (same _name, :54)."""

    _fn = "jwc_wpt"
    _cls = "WaveletPacketTransform"

    def __init__(self, wavelet, context=None):
        super().__init__(wavelet, context)
        self._name = "Wavelet Packet Transform"


class AncientEgyptianDecomposition(BasicTransform):
    """transforms/AncientEgyptianDecomposition.java:42 -- arbitrary-length wrapper: the signal is cut into blocks of
    descending powers of two (tools/MathToolKit.java:57-84), each block goes through the wrapped transform at full
    depth and lands at its own position.  The wrapped transform must be one of the CUDA pyramid transforms; all blocks
    of all signals of a batch are transformed on the device without gathering them (jwc_{fwt,wpt}_aed_*)."""

    def __init__(self, basicTransform, initialWaveletSpaceSize=0):
        if not isinstance(basicTransform, _CudaPyramidBase):
            raise JWaveFailure("AncientEgyptianDecomposition - the CUDA path wraps CudaFastWaveletTransform or "
                               "CudaWaveletPacketTransform")
        self._basicTransform = basicTransform
        self._initialWaveletSpaceSize = initialWaveletSpaceSize
        self._name = basicTransform.getName()

    @staticmethod
    def decompose(number):
        """MathToolKit.decompose: 42 -> [5, 3, 1]."""
        if number < 1:
            raise JWaveFailure("the supported number for decomposition is smaller than one")
        out = []
        while number >= 1:
            p = int(number).bit_length() - 1
            out.append(p)
            number -= 1 << p
        return out

    def _run(self, direction, mat, flags, out):
        t = self._basicTransform
        X = _as_f64(mat)
        B, N = X.shape
        out = _out_buffer(out, X.shape)
        if N == 0 or B == 0:
            return out
        w = t._wavelet
        f0, f1 = ((w.getScalingDeComposition(), w.getWaveletDeComposition()) if direction == "forward"
                  else (w.getScalingReConstruction(), w.getWaveletReConstruction()))
        lib = _native.load()
        f0, f1 = _as_f64(f0), _as_f64(f1)
        fn = getattr(lib, "%s_aed_%s" % (t._fn, direction if direction == "forward" else "inverse"))
        rc = fn(t._context().handle, X.ctypes.data, out.ctypes.data, B, N, _ptr(f0), _ptr(f1), len(f0), flags)
        if rc != 0:
            raise RuntimeError("aed %s failed (%d): %s" % (direction, rc, _native.last_error()))
        return out

    def forward(self, arrTime, flags=0):
        return self._run("forward", _as_f64(arrTime)[None, :], flags, None)[0]

    def reverse(self, arrHilb, flags=0):
        return self._run("reverse", _as_f64(arrHilb)[None, :], flags, None)[0]

    def forwardBatch(self, matTime, flags=0, out=None):
        return self._run("forward", matTime, flags, out)

    def reverseBatch(self, matHilb, flags=0, out=None):
        return self._run("reverse", matHilb, flags, out)


class ArrayView:
    """transforms/EfficientMODWTTransform.java:88-117 -- read-only window on a backing array, no copy."""

    def __init__(self, array, offset, length):
        self._array, self._offset, self._length = array, offset, length

    def get(self, index):
        if index < 0 or index >= self._length:
            raise IndexError("Index: %d, Length: %d" % (index, self._length))
        return float(self._array[self._offset + index])

    def length(self):
        return self._length

    def toArray(self):
        return np.array(self._array[self._offset:self._offset + self._length])


class MODWTCoefficients:
    """transforms/EfficientMODWTTransform.java:28-86 -- the single-backing-array coefficient format
    [W_1 | ... | W_J | V_J].  It is exactly the row layout the device writes (and what forward(double[], level)
    returns, MODWTTransform.java:406-416), so wrapping a GPU result costs nothing."""

    def __init__(self, backingArray, signalLength, levels):
        self._backing = np.asarray(backingArray, dtype=np.float64).reshape(-1)
        if self._backing.size != (levels + 1) * signalLength:
            raise IllegalArgumentException("backing array length %d != (levels + 1) * signalLength"
                                           % self._backing.size)
        self._n, self._levels = signalLength, levels

    def getDetails(self, level):
        if level < 1 or level > self._levels:
            raise IllegalArgumentException("Invalid level: %d" % level)
        off = (level - 1) * self._n
        return np.array(self._backing[off:off + self._n])

    def getApproximation(self):
        off = self._levels * self._n
        return np.array(self._backing[off:off + self._n])

    def getView(self, level):
        if level < 1 or level > self._levels + 1:
            raise IllegalArgumentException("Invalid level: %d" % level)
        return ArrayView(self._backing, (level - 1) * self._n, self._n)

    def getTotalSize(self):
        return int(self._backing.size)

    def backingArray(self):
        return self._backing


class ConvolutionMethod:
    """MODWTTransform.ConvolutionMethod (MODWTTransform.java:148-153).  The reference switches its CPU loops between the
    direct O(N M) and the FFT O(N log N) circular convolution; the device path has one arithmetic (the direct sum, in
    the reference's summation order), so the setting is kept for callers that read it back and changes nothing."""
    AUTO = "AUTO"
    DIRECT = "DIRECT"
    FFT = "FFT"


class CudaMODWTTransform(WaveletTransform):
    """Drop-in for transforms/MODWTTransform.java."""

    MAX_DECOMPOSITION_LEVEL = 13  # MODWTTransform.java:111
    ConvolutionMethod = ConvolutionMethod

    def __init__(self, wavelet, fftThreshold=None, context=None):
        # MODWTTransform.java:180-195: (wavelet) and (wavelet, fftThreshold); the threshold only steers the reference's
        # AUTO choice between its two CPU convolutions and is recorded, not used
        if fftThreshold is not None and not isinstance(fftThreshold, (int, np.integer)):
            context, fftThreshold = fftThreshold, None      # CudaMODWTTransform(wavelet, context)
        super().__init__(wavelet, context)
        self._name = "MODWT"
        self._lock = threading.Lock()
        self._g = None
        self._h = None
        self._fftThreshold = 4096 if fftThreshold is None else int(fftThreshold)   # MODWTTransform.java:128
        self._convolutionMethod = ConvolutionMethod.AUTO

    def setConvolutionMethod(self, method):
        # MODWTTransform.java:202-204
        if method not in (ConvolutionMethod.AUTO, ConvolutionMethod.DIRECT, ConvolutionMethod.FFT):
            raise IllegalArgumentException("unknown convolution method %r" % (method,))
        self._convolutionMethod = method

    def getConvolutionMethod(self):
        # MODWTTransform.java:211-213
        return self._convolutionMethod

    @staticmethod
    def getMaxDecompositionLevel():
        return CudaMODWTTransform.MAX_DECOMPOSITION_LEVEL

    # ---- filter preparation (MODWTTransform.java:452-484, normalize :599-606) --------------------------------
    def initializeFilterCache(self):
        if self._g is None:
            with self._lock:
                if self._g is None:
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
                    h = normalize(self._wavelet.getWaveletDeComposition()) / sf
                    self._g = normalize(self._wavelet.getScalingDeComposition()) / sf
                    self._h = h

    def clearFilterCache(self):
        # MODWTTransform.java:556-569: the GPU path keeps no per-level upsampled filters (the stride is implicit in
        # the kernels); only the two base filters are cached.
        with self._lock:
            self._g = None
            self._h = None

    def precomputeFilters(self, maxLevel):
        # MODWTTransform.java:578-590
        if maxLevel > self.MAX_DECOMPOSITION_LEVEL:
            raise IllegalArgumentException("MODWTTransform#precomputeFilters - maximum supported decomposition level "
                                           "is %d, requested: %d" % (self.MAX_DECOMPOSITION_LEVEL, maxLevel))
        self.initializeFilterCache()

    def _filters(self):
        self.initializeFilterCache()
        g, h = self._g, self._h
        if g is None or h is None:      # cleared by another thread between the two statements
            return self._filters()
        return g, h

    # ---- forwardMODWT / inverseMODWT (MODWTTransform.java:256-306, 337-375) ----------------------------------
    def _check_levels(self, maxLevel, N):
        if maxLevel < 1:
            raise IllegalArgumentException("MODWTTransform#forwardMODWT - decomposition level must be at least 1, "
                                           "requested: %d" % maxLevel)
        if maxLevel > self.MAX_DECOMPOSITION_LEVEL:
            raise IllegalArgumentException("MODWTTransform#forwardMODWT - maximum supported decomposition level is "
                                           "%d, requested: %d" % (self.MAX_DECOMPOSITION_LEVEL, maxLevel))
        if N is None or N == 0:
            return False
        theoretical = int(N).bit_length() - 1
        if maxLevel > theoretical:
            raise IllegalArgumentException("Decomposition level %d exceeds theoretical limit %d for signal length %d"
                                           % (maxLevel, theoretical, N))
        return True

    def _check_inverse_levels(self, maxLevel):
        # the reference's inverse reaches the same limit through getCachedGFilter -> upsample()
        # (MODWTTransform.java:490-515, 618-620): level > MAX_DECOMPOSITION_LEVEL -> IllegalArgumentException
        if maxLevel > self.MAX_DECOMPOSITION_LEVEL:
            raise IllegalArgumentException("MODWTTransform#upsample - maximum supported decomposition level is %d, "
                                           "requested: %d"
                                           % (self.MAX_DECOMPOSITION_LEVEL, maxLevel))

    def forwardMODWT(self, data, maxLevel, flags=0):
        N = 0 if data is None else len(data)
        if not self._check_levels(maxLevel, N):
            return [np.empty(0) for _ in range(maxLevel + 1)]
        x = _as_f64(data)
        g, h = self._filters()
        out = np.empty((maxLevel + 1, N))
        self._call("jwc_modwt_forward", x, out, 1, N, maxLevel, g, h, flags)
        return out

    def inverseMODWT(self, coefficients, flags=0):
        if coefficients is None or len(coefficients) == 0:
            return np.empty(0)
        maxLevel = len(coefficients) - 1
        if maxLevel <= 0:
            return np.empty(0)
        c = _as_f64(coefficients)
        N = c.shape[1]
        if N == 0:
            return np.empty(0)
        self._check_inverse_levels(maxLevel)
        g, h = self._filters()
        x = np.empty(N)
        self._call("jwc_modwt_inverse", c, x, 1, N, maxLevel, g, h, flags)
        return x

    # ---- flattened 1-D interface (MODWTTransform.java:389-443, 854-912) --------------------------------------
    def forward(self, arrTime, level=None):
        if arrTime is None or len(arrTime) == 0:
            return np.empty(0)
        if level is None:
            level = self.calcExponent(len(arrTime))       # :858 (throws JWaveFailure for non-2^p)
            return self.forwardMODWT(arrTime, level).reshape(-1)
        if not self.isBinary(len(arrTime)):
            raise JWaveFailure("MODWTTransform#forward - given array length is not 2^p | p E N ... = "
                               "1, 2, 4, 8, 16, 32, .. ")
        maxLevel = self.calcExponent(len(arrTime))
        if level < 0 or level > maxLevel:
            raise JWaveFailure("MODWTTransform#forward - given level is out of range for given array")
        if level > self.MAX_DECOMPOSITION_LEVEL:
            raise JWaveFailure("MODWTTransform#forward - maximum supported decomposition level is %d, requested: %d"
                               % (self.MAX_DECOMPOSITION_LEVEL, level))
        return self.forwardMODWT(arrTime, level).reshape(-1)

    def reverse(self, arrHilb, level=None):
        if arrHilb is None or len(arrHilb) == 0:
            return np.empty(0)
        total = len(arrHilb)
        if level is None:
            # :888-897 -- smallest 2^p N with total % N == 0 and total/N - 1 <= log2 N
            N = 0
            levels = 0
            for testN in range(1, total + 1):
                if total % testN == 0:
                    testLevels = total // testN - 1
                    if testLevels >= 0 and self.isBinary(testN) and testLevels <= self.calcExponent(testN):
                        N, levels = testN, testLevels
                        break
            if N == 0:
                raise JWaveFailure("MODWTTransform#reverse - Invalid flattened coefficient array length. "
                                   "Cannot determine original signal dimensions.")
        else:
            levels = level
            N = total // (level + 1)
            if not self.isBinary(N):
                raise JWaveFailure("MODWTTransform#reverse - Invalid coefficient array for given level")
            if total != N * (level + 1):
                raise JWaveFailure("MODWTTransform#reverse - Coefficient array length does not match expected size "
                                   "for given level")
        return self.inverseMODWT(_as_f64(arrHilb).reshape(levels + 1, N))

    def forwardMODWTWindows(self, series, window, hop, maxLevel, flags=0, out=None):
        """Sliding-window analysis (MODWTSlidingWindowTest.java:20-70): forwardMODWT of every window
        series[w*hop : w*hop + window]; returns [nwin][maxLevel+1][window].  One device pass over the series; the
        windows are read in place."""
        x = _as_f64(series)
        if x.ndim != 1 or window < 1 or hop < 1 or len(x) < window:
            raise IllegalArgumentException("need a 1-D series with 1 <= window <= len(series) and hop >= 1")
        self._check_levels(maxLevel, window)
        nwin = (len(x) - window) // hop + 1
        out = _out_buffer(out, (nwin, maxLevel + 1, window))
        g, h = self._filters()
        lib = _native.load()
        g, h = _as_f64(g), _as_f64(h)
        rc = lib.jwc_modwt_forward_windows(self._context().handle, x.ctypes.data, out.ctypes.data, len(x), window, hop,
                                           maxLevel, _ptr(g), _ptr(h), len(g), flags)
        if rc != 0:
            raise RuntimeError("jwc_modwt_forward_windows failed (%d): %s" % (rc, _native.last_error()))
        return out

    def forwardMODWTWindowsCompressDevice(self, d_series, d_coeffs, series_len, window, hop, maxLevel, threshold,
                                          d_magnitude, stream=0, flags=0, slot=0):
        """The reference's compression chain on device buffers: forwardMODWT of every window (as forwardMODWTWindows),
        then CompressorMagnitude(threshold).compress over all coefficients, in place in d_coeffs
        ([nwin][maxLevel+1][window]); the sum of |c| is taken in the transform's store epilogue, mean |c| lands in the
        device double d_magnitude.  Arguments are raw device addresses."""
        self._check_levels(maxLevel, window)
        g, h = self._filters()
        lib = _native.load()
        g, h = _as_f64(g), _as_f64(h)
        rc = lib.jwc_modwt_forward_windows_compress_dev(
            self._context().handle, slot, ctypes.c_void_p(stream if stream else 1), ctypes.c_void_p(d_series),
            ctypes.c_void_p(d_coeffs), series_len, window, hop, maxLevel, _ptr(g), _ptr(h), len(g), flags,
            float(threshold), ctypes.c_void_p(d_magnitude))
        if rc != 0:
            raise RuntimeError("jwc_modwt_forward_windows_compress_dev failed (%d): %s" % (rc, _native.last_error()))

    def forwardMODWTCoefficients(self, data, maxLevel, flags=0):
        """The result of forwardMODWT in the MODWTCoefficients wire format (one backing array, level views).  Rows
        carry forwardMODWT's meaning (W_j = h~ conv V, V_J last); the reference's forwardMODWTEfficient
        (EfficientMODWTTransform.java:151-170) labels the two branches the other way round and is covered by none of
        its tests, so it is not reproduced."""
        c = self.forwardMODWT(data, maxLevel, flags=flags)
        return MODWTCoefficients(c.reshape(-1), c.shape[1], maxLevel)

    def inverseMODWTCoefficients(self, coeffs, flags=0):
        return self.inverseMODWT(coeffs.backingArray().reshape(coeffs._levels + 1, coeffs._n), flags=flags)

    # ---- batched host-buffer entry points ------------------------------------------------------------------
    def forwardMODWTBatch(self, matTime, maxLevel, flags=0, out=None):
        X = _as_f64(matTime)
        B, N = X.shape
        self._check_levels(maxLevel, N)
        g, h = self._filters()
        out = _out_buffer(out, (B, maxLevel + 1, N))
        self._call("jwc_modwt_forward", X, out, B, N, maxLevel, g, h, flags)
        return out

    def inverseMODWTBatch(self, coeffs, flags=0, out=None):
        C = _as_f64(coeffs)
        B, J1, N = C.shape
        self._check_inverse_levels(J1 - 1)
        g, h = self._filters()
        out = _out_buffer(out, (B, N))
        self._call("jwc_modwt_inverse", C, out, B, N, J1 - 1, g, h, flags)
        return out

    # ---- one long series split over the context's devices (halo exchange between ring neighbours) -----------------
    def _split(self, fn_name, src_ptrs, dst_ptrs, n, maxLevel, flags):
        lib = _native.load()
        g, h = self._filters()
        P = self._context().num_devices()
        if len(src_ptrs) != P or len(dst_ptrs) != P:
            raise ValueError("need one chunk pointer per device of the context (%d)" % P)
        a = (ctypes.c_void_p * P)(*src_ptrs)
        b = (ctypes.c_void_p * P)(*dst_ptrs)
        rc = getattr(lib, fn_name)(self._context().handle, a, b, n, maxLevel, _ptr(g), _ptr(h), len(g), flags)
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (fn_name, rc, _native.last_error()))

    def forwardMODWTSplitDevice(self, d_x_chunks, d_coeff_chunks, n, maxLevel, flags=0):
        """d_x_chunks[p]: device address of x[n*p/P .. n*(p+1)/P) on device p; d_coeff_chunks[p]: [J+1][chunk_len]."""
        self._check_levels(maxLevel, n)
        self._split("jwc_modwt_forward_split_dev", d_x_chunks, d_coeff_chunks, n, maxLevel, flags)

    def inverseMODWTSplitDevice(self, d_coeff_chunks, d_x_chunks, n, maxLevel, flags=0):
        self._split("jwc_modwt_inverse_split_dev", d_coeff_chunks, d_x_chunks, n, maxLevel, flags)

    # ---- device-resident ----------------------------------------------------------------------------------------
    def forwardMODWTDevice(self, d_x, d_coeffs, batch, n, maxLevel, stream=0, flags=0, slot=0):
        self._check_levels(maxLevel, n)
        g, h = self._filters()
        self._call_dev("jwc_modwt_forward", d_x, d_coeffs, batch, n, maxLevel, g, h, flags, stream, slot)

    def inverseMODWTDevice(self, d_coeffs, d_x, batch, n, maxLevel, stream=0, flags=0, slot=0):
        self._check_inverse_levels(maxLevel)
        g, h = self._filters()
        self._call_dev("jwc_modwt_inverse", d_coeffs, d_x, batch, n, maxLevel, g, h, flags, stream, slot)
