"""CPU-only: the C-ABI library loads and exports every symbol include/jwavecuda.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "jwavecuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"JWC_API\s+[\w\s\*]+?\b(jwc_\w+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = _header_symbols()
    for t in ("modwt", "fwt", "wpt"):
        for d in ("forward", "inverse"):
            assert "jwc_%s_%s" % (t, d) in syms
            assert "jwc_%s_%s_dev" % (t, d) in syms
    assert "jwc_create" in syms and "jwc_last_error" in syms


def test_library_builds_and_exports_every_declared_symbol(jw):
    from jwave_pro_b200 import _native
    import importlib.util
    spec = importlib.util.spec_from_file_location("jwc_build", os.path.join(ROOT, "jwave-pro_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    so = mod.build()
    assert os.path.exists(so)
    lib = ctypes.CDLL(so)
    declared = _header_symbols()
    assert sorted(_native.SYMBOLS) == declared, "Python binding list and header disagree"
    for s in declared:
        assert hasattr(lib, s), "missing export %s" % s
    assert b"sm_100a" in ctypes.cast(lib.jwc_version, ctypes.CFUNCTYPE(ctypes.c_char_p))()


def test_sass_is_sm_100a_only():
    import subprocess
    so = os.path.join(ROOT, "jwave-pro_b200", "libjwavecuda.so")
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_device(jw):
    """On a box without a GPU the product must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the loud-failure path is exercised on the CPU box")
    with pytest.raises(jw.NativeLibraryError):
        jw.CudaMODWTTransform(jw.wavelets.Haar1()).forwardMODWT([1.0, 2.0, 3.0, 4.0], 1)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under jwave-pro_b200/ may import, link or call it."""
    pkg = os.path.join(ROOT, "jwave-pro_b200")
    for dp, _, files in os.walk(pkg):
        if os.path.basename(dp) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in txt.lower(), "%s mentions the oracle" % f
