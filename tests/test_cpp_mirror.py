"""The C++ host mirror (include/jwave_cuda.hpp): compiles everywhere (CPU test), runs on the GPU box (gpu test)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_mirror")


def _build():
    so_dir = os.path.join(ROOT, "jwave-pro_b200")
    src = os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp")
    deps = [src, os.path.join(ROOT, "include", "jwave_cuda.hpp"), os.path.join(ROOT, "include", "jwavecuda.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", EXE, src, "-L" + so_dir, "-ljwavecuda",
                               "-Wl,-rpath," + so_dir])
    return EXE


def test_cpp_mirror_compiles_and_links(jw):
    assert os.path.exists(_build())


@pytest.mark.gpu
def test_cpp_mirror_runs(gpu_ctx):
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host mirror ok" in r.stdout
