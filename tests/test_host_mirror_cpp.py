"""Builds tests/cpp/test_host_mirror.cpp against include/ellalgo_cuda.hpp + libellcuda.so and runs it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ellalgo_rs_b200", "csrc")
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_host_mirror")


def build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    from ellalgo_rs_b200 import _lib
    _lib.build()
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp"), "-o", EXE, "-L", CSRC, "-lellcuda",
                    f"-Wl,-rpath,{CSRC}"], check=True)


def test_cpp_host_mirror_builds_and_fails_loudly_without_gpu():
    build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-fallback check needs a CPU-only box")
    p = subprocess.run([EXE, "--no-gpu"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr


@pytest.mark.gpu
def test_cpp_host_mirror_reference_tests_on_gpu():
    build()
    p = subprocess.run([EXE], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
