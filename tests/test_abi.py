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


def test_java_ffm_descriptors_match_the_header():
    """The Java drop-in cannot be compiled here (no JDK), so its Panama FunctionDescriptors are checked statically: every
    symbol java/.../JwcNative.java binds must be declared in include/jwavecuda.h with the same return type and the
    same argument kinds in the same order (pointer -> ADDRESS, int64_t / size_t -> JAVA_LONG, int / unsigned -> JAVA_INT)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "jwavecuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"JWC_API\s+([\w\s\*]+?)\s*\b(jwc_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3)

        def kind(t):
            t = t.strip()
            if "*" in t:
                return "ADDRESS"
            base = re.sub(r"\b(const|unsigned|signed)\b", "", t).split()
            word = base[0] if base else "int"           # "unsigned flags" -> int-sized
            if "int64_t" in t or "size_t" in t or "uint64_t" in t:
                return "JAVA_LONG"
            if word in ("int", "flags") or t.startswith("unsigned") or "int" in t.split():
                return "JAVA_INT"
            if word == "double":
                return "JAVA_DOUBLE"
            if word == "void":
                return "VOID"
            raise AssertionError("unmapped C type %r in %s" % (t, name))
        argl = [] if args.strip() in ("", "void") else [kind(a.rsplit(None, 1)[0] if "*" not in a.rsplit(None, 1)[-1] else a)
                                                         for a in args.split(",")]
        protos[name] = [kind(ret)] + argl
    java = open(os.path.join(root, "java", "jwave", "transforms", "cuda", "JwcNative.java")).read()
    java = re.sub(r"//[^\n]*", " ", java)

    def parse_fd(text):
        m = re.match(r"FunctionDescriptor\.(of|ofVoid)\s*\((.*)\)\s*$", text.strip(), flags=re.S)
        assert m, text
        parts = [p.strip() for p in m.group(2).split(",") if p.strip()]
        return (["VOID"] + parts) if m.group(1) == "ofVoid" else parts

    def balanced(s, start):   # text of the call whose '(' is at or after `start`, up to its matching ')'
        i = s.index("(", start)
        depth, j = 0, i
        while True:
            depth += s[j] == "("
            depth -= s[j] == ")"
            if depth == 0:
                return s[start:j + 1]
            j += 1
    variables = {m.group(1): parse_fd(balanced(java, m.start(2)))
                 for m in re.finditer(r"FunctionDescriptor\s+(\w+)\s*=\s*(FunctionDescriptor\.)", java)}
    arrays = {m.group(1): re.findall(r'"(jwc_\w+)"', m.group(2))
              for m in re.finditer(r"String\[\]\s+(\w+)\s*=\s*\{(.*?)\}", java, flags=re.S)}
    bound = {}
    for m in re.finditer(r'handle\(\s*"(jwc_\w+)"\s*,\s*(FunctionDescriptor\.)', java):
        bound[m.group(1)] = parse_fd(balanced(java, m.start(2)))
    for m in re.finditer(r"handle\(\s*(\w+)\[i\]\s*,\s*(\w+)\s*\)", java):
        for name in arrays[m.group(1)]:
            bound[name] = variables[m.group(2)]
    assert len(bound) >= 20, sorted(bound)
    for name, fd in sorted(bound.items()):
        assert name in protos, "%s is bound in JwcNative.java but not declared in jwavecuda.h" % name
        assert fd == protos[name], (name, fd, protos[name])
