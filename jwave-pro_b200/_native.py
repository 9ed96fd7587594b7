"""ctypes binding of libjwavecuda.so (include/jwavecuda.h).  Fails loudly: no library or no GPU => NativeLibraryError.

This is the same C ABI the Java classes bind through Panama FFM (java/jwave/transforms/cuda/JwcNative.java).
"""
import ctypes
import os
import threading

from .exceptions import NativeLibraryError

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("JWAVECUDA_LIB", os.path.join(_HERE, "libjwavecuda.so"))

_dp = ctypes.POINTER(ctypes.c_double)
_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_u32 = ctypes.c_uint

FLAG_EXACT = 1
FLAG_FORCE_GENERIC = 2
MAX_TAPS = 64

# every exported symbol of include/jwavecuda.h (tests/test_abi.py checks the .so against this list and the header)
_TRANSFORMS = ["modwt_forward", "modwt_inverse", "fwt_forward", "fwt_inverse", "wpt_forward", "wpt_inverse"]
_TRANSFORMS_AED = ["fwt_aed_forward", "fwt_aed_inverse", "wpt_aed_forward", "wpt_aed_inverse"]
_TRANSFORMS_2D = ["fwt2d_forward", "fwt2d_inverse", "wpt2d_forward", "wpt2d_inverse"]
_TRANSFORMS_3D = ["fwt3d_forward", "fwt3d_inverse", "wpt3d_forward", "wpt3d_inverse"]
SYMBOLS = (["jwc_create", "jwc_destroy", "jwc_num_devices", "jwc_device_ordinal", "jwc_last_error", "jwc_version",
            "jwc_launch_count", "jwc_set_tuning", "jwc_get_tuning", "jwc_alloc_pinned", "jwc_free_pinned",
            "jwc_alloc_device", "jwc_free_device", "jwc_copy_to_device", "jwc_copy_to_host", "jwc_synchronize",
            "jwc_release_scratch"]
           + ["jwc_" + t for t in _TRANSFORMS] + ["jwc_" + t + "_dev" for t in _TRANSFORMS]
           + ["jwc_modwt_forward_split_dev", "jwc_modwt_inverse_split_dev"]
           + ["jwc_fwt_forward_split_dev", "jwc_fwt_inverse_split_dev", "jwc_wpt_forward_split_dev",
              "jwc_wpt_inverse_split_dev", "jwc_dwt_split_levels"]
           + ["jwc_modwt_forward_windows", "jwc_modwt_forward_windows_dev", "jwc_compress_magnitude",
              "jwc_compress_magnitude_dev", "jwc_modwt_forward_windows_compress_dev"]
           + ["jwc_diag_dfma_tflops", "jwc_diag_copy_gbs"]
           + ["jwc_" + t for t in _TRANSFORMS_2D] + ["jwc_" + t + "_dev" for t in _TRANSFORMS_2D]
           + ["jwc_" + t for t in _TRANSFORMS_3D] + ["jwc_" + t + "_dev" for t in _TRANSFORMS_3D]
           + ["jwc_" + t for t in _TRANSFORMS_AED] + ["jwc_" + t + "_dev" for t in _TRANSFORMS_AED])

_lib = None
_lock = threading.Lock()


def load():
    """dlopen the library and declare signatures (no CUDA call is made here)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(SO_PATH):
            raise NativeLibraryError(
                "libjwavecuda.so not found at %s -- build it with `python jwave-pro_b200/build.py` "
                "(there is no CPU fallback)" % SO_PATH)
        try:
            lib = ctypes.CDLL(SO_PATH)
        except OSError as e:
            raise NativeLibraryError("cannot load %s: %s" % (SO_PATH, e))
        lib.jwc_create.argtypes = [ctypes.POINTER(_int), _int]
        lib.jwc_create.restype = _vp
        lib.jwc_destroy.argtypes = [_vp]
        lib.jwc_destroy.restype = None
        lib.jwc_num_devices.argtypes = [_vp]
        lib.jwc_num_devices.restype = _int
        lib.jwc_device_ordinal.argtypes = [_vp, _int]
        lib.jwc_device_ordinal.restype = _int
        lib.jwc_last_error.argtypes = []
        lib.jwc_last_error.restype = ctypes.c_char_p
        lib.jwc_version.argtypes = []
        lib.jwc_version.restype = ctypes.c_char_p
        lib.jwc_launch_count.argtypes = [_vp]
        lib.jwc_launch_count.restype = ctypes.c_uint64
        lib.jwc_set_tuning.argtypes = [_vp, ctypes.c_char_p, _int]
        lib.jwc_set_tuning.restype = _int
        lib.jwc_get_tuning.argtypes = [_vp, ctypes.c_char_p, ctypes.POINTER(_int)]
        lib.jwc_get_tuning.restype = _int
        lib.jwc_alloc_pinned.argtypes = [ctypes.c_size_t]
        lib.jwc_alloc_pinned.restype = _vp
        lib.jwc_free_pinned.argtypes = [_vp]
        lib.jwc_free_pinned.restype = None
        lib.jwc_alloc_device.argtypes = [_vp, _int, ctypes.c_size_t]
        lib.jwc_alloc_device.restype = _vp
        lib.jwc_free_device.argtypes = [_vp, _int, _vp]
        lib.jwc_free_device.restype = None
        lib.jwc_copy_to_device.argtypes = [_vp, _int, _vp, _vp, ctypes.c_size_t]
        lib.jwc_copy_to_device.restype = _int
        lib.jwc_copy_to_host.argtypes = [_vp, _int, _vp, _vp, ctypes.c_size_t]
        lib.jwc_copy_to_host.restype = _int
        lib.jwc_synchronize.argtypes = [_vp]
        lib.jwc_synchronize.restype = _int
        lib.jwc_release_scratch.argtypes = [_vp]
        lib.jwc_release_scratch.restype = _int
        for t in _TRANSFORMS:
            fn = getattr(lib, "jwc_" + t)
            fn.argtypes = [_vp, _vp, _vp, _i64, _i64, _int, _dp, _dp, _int, _u32]
            fn.restype = _int
            fn = getattr(lib, "jwc_" + t + "_dev")
            fn.argtypes = [_vp, _int, _vp, _vp, _vp, _i64, _i64, _int, _dp, _dp, _int, _u32]
            fn.restype = _int
        lib.jwc_modwt_forward_windows.argtypes = [_vp, _vp, _vp, _i64, _i64, _i64, _int, _dp, _dp, _int, _u32]
        lib.jwc_modwt_forward_windows.restype = _int
        lib.jwc_modwt_forward_windows_dev.argtypes = [_vp, _int, _vp, _vp, _vp, _i64, _i64, _i64, _int, _dp, _dp, _int,
                                                      _u32]
        lib.jwc_modwt_forward_windows_dev.restype = _int
        lib.jwc_modwt_forward_windows_compress_dev.argtypes = [_vp, _int, _vp, _vp, _vp, _i64, _i64, _i64, _int, _dp, _dp,
                                                               _int, _u32, ctypes.c_double, _vp]
        lib.jwc_modwt_forward_windows_compress_dev.restype = _int
        lib.jwc_compress_magnitude.argtypes = [_vp, _vp, _vp, _i64, ctypes.c_double, _dp]
        lib.jwc_compress_magnitude.restype = _int
        lib.jwc_compress_magnitude_dev.argtypes = [_vp, _int, _vp, _vp, _vp, _i64, ctypes.c_double, _vp]
        lib.jwc_compress_magnitude_dev.restype = _int
        for t in _TRANSFORMS_AED:
            fn = getattr(lib, "jwc_" + t)
            fn.argtypes = [_vp, _vp, _vp, _i64, _i64, _dp, _dp, _int, _u32]
            fn.restype = _int
            fn = getattr(lib, "jwc_" + t + "_dev")
            fn.argtypes = [_vp, _int, _vp, _vp, _vp, _i64, _i64, _dp, _dp, _int, _u32]
            fn.restype = _int
        for t in _TRANSFORMS_2D:
            fn = getattr(lib, "jwc_" + t)
            fn.argtypes = [_vp, _vp, _vp, _i64, _i64, _i64, _int, _int, _dp, _dp, _int, _u32]
            fn.restype = _int
            fn = getattr(lib, "jwc_" + t + "_dev")
            fn.argtypes = [_vp, _int, _vp, _vp, _vp, _i64, _i64, _i64, _int, _int, _dp, _dp, _int, _u32]
            fn.restype = _int
        for t in _TRANSFORMS_3D:
            fn = getattr(lib, "jwc_" + t)
            fn.argtypes = [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _int, _int, _int, _dp, _dp, _int, _u32]
            fn.restype = _int
            fn = getattr(lib, "jwc_" + t + "_dev")
            fn.argtypes = [_vp, _int, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _int, _int, _int, _dp, _dp, _int, _u32]
            fn.restype = _int
        lib.jwc_diag_dfma_tflops.argtypes = [_vp, _int, _dp]
        lib.jwc_diag_dfma_tflops.restype = _int
        lib.jwc_diag_copy_gbs.argtypes = [_vp, _int, ctypes.c_size_t, _dp]
        lib.jwc_diag_copy_gbs.restype = _int
        lib.jwc_dwt_split_levels.argtypes = [_vp, _i64, _int]
        lib.jwc_dwt_split_levels.restype = _int
        for nm in ("jwc_modwt_forward_split_dev", "jwc_modwt_inverse_split_dev", "jwc_fwt_forward_split_dev",
                   "jwc_fwt_inverse_split_dev", "jwc_wpt_forward_split_dev", "jwc_wpt_inverse_split_dev"):
            fn = getattr(lib, nm)
            fn.argtypes = [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i64, _int, _dp, _dp, _int, _u32]
            fn.restype = _int
        _lib = lib
        return _lib


def last_error():
    return load().jwc_last_error().decode("utf-8", "replace")


class Context:
    """Owns one jwc_ctx.  devices=None -> current CUDA device."""

    def __init__(self, devices=None):
        lib = load()
        if devices is None:
            h = lib.jwc_create(None, 0)
        else:
            arr = (_int * len(devices))(*devices)
            h = lib.jwc_create(arr, len(devices))
        if not h:
            raise NativeLibraryError("jwc_create failed: %s" % last_error())
        self._h = h
        self._lib = lib

    @property
    def handle(self):
        return self._h

    def num_devices(self):
        return self._lib.jwc_num_devices(self._h)

    def launch_count(self):
        return int(self._lib.jwc_launch_count(self._h))

    def set_tuning(self, key, value):
        rc = self._lib.jwc_set_tuning(self._h, key.encode(), int(value))
        if rc != 0:
            raise ValueError(last_error())

    def dfma_tflops(self, slot=0):
        """fp64 FMA rate of device `slot` measured now (TFLOP/s): the second roofline of the fp64-bound configs."""
        v = ctypes.c_double(0.0)
        rc = self._lib.jwc_diag_dfma_tflops(self._h, slot, ctypes.byref(v))
        if rc != 0:
            raise RuntimeError(last_error())
        return v.value

    def copy_gbs(self, nbytes=1 << 30, slot=0):
        """plain device copy rate of device `slot` measured now (read + write GB/s)."""
        v = ctypes.c_double(0.0)
        rc = self._lib.jwc_diag_copy_gbs(self._h, slot, nbytes, ctypes.byref(v))
        if rc != 0:
            raise RuntimeError(last_error())
        return v.value

    def synchronize(self):
        rc = self._lib.jwc_synchronize(self._h)
        if rc != 0:
            raise RuntimeError(last_error())

    def release_scratch(self):
        rc = self._lib.jwc_release_scratch(self._h)
        if rc != 0:
            raise RuntimeError(last_error())

    def close(self):
        if self._h:
            self._lib.jwc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx
