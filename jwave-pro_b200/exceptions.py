"""Error types mirroring jwave/exceptions/*.java.

The reference's JWaveException extends Throwable (exceptions/JWaveException.java:32) and JWaveFailure /
JWaveError extend it; shape errors are JWaveFailure, MODWT level errors are the *unchecked*
IllegalArgumentException (transforms/MODWTTransform.java:257-282).  Python has no checked exceptions, so the
mirror keeps the hierarchy and the message substrings the reference's tests grep for.
"""


class JWaveException(Exception):
    pass


class JWaveFailure(JWaveException):
    pass


class JWaveError(JWaveException):
    pass


class IllegalArgumentException(ValueError):
    """Stand-in for java.lang.IllegalArgumentException thrown by MODWTTransform.forwardMODWT."""


class NativeLibraryError(RuntimeError):
    """libjwavecuda.so missing / not loadable / no CUDA device.  There is no CPU fallback."""
