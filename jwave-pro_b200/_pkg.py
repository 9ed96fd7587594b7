"""Public surface of the package (imported by the `jwave_pro_b200` shim and by __init__)."""
from . import wavelets
from ._native import FLAG_EXACT, FLAG_FORCE_GENERIC, Context, default_context
from .compressions import CompressorMagnitude
from .exceptions import (IllegalArgumentException, JWaveError, JWaveException, JWaveFailure, NativeLibraryError)
from .transforms import (AncientEgyptianDecomposition, ArrayView, BasicTransform, ConvolutionMethod, CudaFastWaveletTransform,
                         CudaMODWTTransform, CudaWaveletPacketTransform, MODWTCoefficients, WaveletTransform)
from .wavelets import Wavelet

__all__ = ["wavelets", "Wavelet", "Context", "default_context", "FLAG_EXACT", "FLAG_FORCE_GENERIC",
           "BasicTransform", "WaveletTransform", "CudaFastWaveletTransform", "CudaWaveletPacketTransform",
           "CudaMODWTTransform", "ConvolutionMethod", "MODWTCoefficients", "ArrayView", "CompressorMagnitude", "AncientEgyptianDecomposition", "JWaveException", "JWaveFailure", "JWaveError", "IllegalArgumentException",
           "NativeLibraryError"]
