"""Host mirror of the reference's magnitude compressor, running on the GPU (jwc_compress_magnitude).

Reference: compressions/Compressor.java:58-230 (threshold handling, select, compression rate) and
compressions/CompressorMagnitude.java:44-140 (magnitude = mean |c| of the whole array / matrix / space).
The step right after a transform in the reference's compression path.  No CPU path.
"""
import ctypes

import numpy as np

from . import _native


class CompressorMagnitude:
    def __init__(self, threshold=1.0, context=None):
        # Compressor.java:65-85: a non-positive threshold is reported and replaced by the default 1.0
        if threshold <= 0.0:
            print("Compressor - given threshold should be larger than zero!")
            print("Compressor - setting threshold to default value: 1.0")
            threshold = 1.0
        self._threshold = float(threshold)
        self._magnitude = 0.0
        self._ctx = context

    def getThreshold(self):
        return self._threshold

    def getMagnitude(self):
        return self._magnitude

    def compress(self, hilb):
        """double[] / double[][] / double[][][] alike: one mean over every value, values below mean * threshold -> 0."""
        x = np.ascontiguousarray(hilb, dtype=np.float64)
        out = np.empty_like(x)
        mag = ctypes.c_double(0.0)
        ctx = self._ctx if self._ctx is not None else _native.default_context()
        rc = _native.load().jwc_compress_magnitude(ctx.handle, x.ctypes.data, out.ctypes.data, x.size, self._threshold,
                                                   ctypes.byref(mag))
        if rc != 0:
            raise RuntimeError("jwc_compress_magnitude failed (%d): %s" % (rc, _native.last_error()))
        self._magnitude = mag.value
        return out

    @staticmethod
    def calcCompressionRate(arr):
        # Compressor.java:182-196: percentage of exact zeros
        a = np.asarray(arr)
        zeros = int(np.count_nonzero(a == 0.0))
        return zeros / a.size * 100.0 if zeros else 0.0
