"""CPU tests of the CTA-per-problem kernel's DEVICE CODE (ellalgo_rs_b200/csrc/ellcta_device.cuh) compiled for the host by
tests/cpp/emu_ellcta.cpp — one OS thread per CUDA thread, pthread barriers for __syncthreads / ballots / shuffles:

* the emulated kernel must reproduce the CPU checker bit for bit (cursor handling, first violator in traversal order across
  warps, packed-triangle indexing, the balanced rank-1, the non-symmetric start) — kernel logic checked without a GPU;
* under ThreadSanitizer, where the emulated threads really run asynchronously, it must be free of data races: every
  shared-memory exchange of the kernel is ordered by a barrier (SURVEY.md §5 "race detection"; compute-sanitizer's racecheck
  is not available on the GPU pool), and the detector itself is checked by dropping one barrier.

The GPU parity tests of the same code are tests/test_ellcta_gpu.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from ellalgo_rs_b200 import workloads as W
from ellalgo_rs_b200.problem import BatchOutput
from parity import assert_parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "emu_ellcta.cpp")
DEPS = [SRC] + [os.path.join(ROOT, "ellalgo_rs_b200", "csrc", f) for f in ("ellcta_device.cuh", "common.cuh", "ell_calc.cuh")]
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")
LIB = os.path.join(BUILD, "libemu_ellcta.so")
TSAN = os.path.join(BUILD, "emu_ellcta_tsan")
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
COMMON = ["g++", "-std=c++17", "-g", "-ffp-contract=off", "-Wno-attributes", "-I", CUDA_INC, SRC]


def _stale(out):
    return not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in DEPS)


def _build(out, extra):
    os.makedirs(BUILD, exist_ok=True)
    if _stale(out):
        subprocess.run(COMMON + extra + ["-o", out, "-lpthread", "-ldl"], check=True, capture_output=True, text=True)
    return out


@pytest.fixture(scope="module")
def emu():
    lib = C.CDLL(_build(LIB, ["-O2", "-fPIC", "-shared"]))
    lib.emu_ellcta_solve.restype = C.c_int
    lib.emu_ellcta_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int]

    def solve(prob, ctas=2):
        out = BatchOutput.empty(prob.count, prob.n)
        d, r = prob.to_desc(0), out.to_result()
        assert lib.emu_ellcta_solve(C.byref(d), C.byref(r), ctas) == 0
        return out
    return solve


def test_emulated_bowl_equals_checker(ref, emu):
    rng = np.random.default_rng(7)
    for n, count, iters in ((40, 5, 120), (64, 3, 80), (70, 3, 80), (100, 2, 60)):
        prob = W.bowl(count, n=n, max_iters=iters)
        assert_parity(emu(prob), ref.solve(prob), exact=True, what=f"emulated bowl n={n}")
    n = 45
    prob = W.bowl(3, n=n, max_iters=70)
    prob.kappa0 = np.ones(1)
    prob.mq0 = np.ascontiguousarray(100.0 * (np.eye(n)[None] + rng.uniform(-1, 1, (3, n, n)) / n))   # non-symmetric start
    for nd in (False, True):
        prob.no_defer_trick = nd
        assert_parity(emu(prob), ref.solve(prob), exact=True, what=f"emulated bowl n={n} matrix start no_defer={nd}")
    prob = W.bowl(3, n=33, max_iters=50)
    prob.kappa0 = np.ones(1)
    prob.diag0 = rng.uniform(50.0, 150.0, (3, 33))
    prob.use_parallel_cut = False
    assert_parity(emu(prob, ctas=1), ref.solve(prob), exact=True, what="emulated bowl n=33 diag start")


def test_emulated_lowpass_equals_checker(ref, emu):
    for n, count, iters in ((48, 2, 700), (70, 2, 300)):
        prob = W.lowpass(count, ref.lowpass_table(n), n=n, max_iters=iters)
        g, r = emu(prob), ref.solve(prob)
        assert_parity(g, r, exact=True, what=f"emulated lowpass n={n}")
    assert g.num_iters.max() == 300
    n = 64                                            # band edges on multiples of the CTA's round length, an empty transition band
    prob = W.lowpass(4, ref.lowpass_table(n), n=n, max_iters=120)
    m1 = 15 * n - 1
    for i, (wp, ws) in enumerate([(0.001, 0.002), (0.12, 0.12), (63.5 / m1, 127.5 / m1), (0.05, 0.995)]):
        prob.params[i, 0], prob.params[i, 1] = wp, ws
    assert_parity(emu(prob), ref.solve(prob), exact=True, what="emulated lowpass band layouts")


def _tsan_exe():
    return _build(TSAN, ["-O1", "-fsanitize=thread", "-DEMU_MAIN"])


def test_no_data_race_under_thread_sanitizer():
    p = subprocess.run([_tsan_exe()], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "ThreadSanitizer" not in p.stderr, p.stdout[-2000:] + p.stderr[-4000:]
    assert "lowpass n=40: rc 0" in p.stdout and "3 of 3 with xbest" in p.stdout     # the shrink path ran too


def test_race_detector_sees_a_dropped_barrier():
    env = dict(os.environ, EMU_DROP_BARRIER="3")
    try:
        p = subprocess.run([_tsan_exe()], capture_output=True, text=True, timeout=120, env=env)
        err = p.stderr
    except subprocess.TimeoutExpired as e:            # without the barrier the threads may also diverge and hang
        err = (e.stderr or b"").decode() if isinstance(e.stderr, bytes) else (e.stderr or "")
    assert "WARNING: ThreadSanitizer: data race" in err
