"""CPU tests of DEVICE CODE compiled for the host (tests/cpp/cuda_emu.h: one OS thread per CUDA thread, pthread barriers for
__syncthreads / ballots / shuffles) — the two kernels whose threads hand data to other warps through shared memory in every
phase: the CTA-per-problem batched `Ell` kernel (csrc/ellcta_device.cuh, tests/cpp/emu_ellcta.cpp) and the single-handle
`EllStable` update for n > 64 (csrc/stable_large_device.cuh, tests/cpp/emu_stable_large.cpp):

* the emulated kernel must reproduce the CPU checker bit for bit (cursor handling, first violator in traversal order across
  warps, packed-triangle indexing, the balanced rank-1, the non-symmetric start) — kernel logic checked without a GPU;
* under ThreadSanitizer, where the emulated threads really run asynchronously, it must be free of data races: every
  shared-memory exchange of the kernel is ordered by a barrier (SURVEY.md §5 "race detection"; compute-sanitizer's racecheck
  is not available on the GPU pool), and the detector itself is checked by dropping one barrier.

The GPU parity tests of the same code are tests/test_ellcta_gpu.py and tests/test_stream_gpu.py::test_ellstable_*."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from ellalgo_rs_b200 import workloads as W
from ellalgo_rs_b200.problem import BatchOutput
from parity import assert_parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
CSRC = os.path.join(ROOT, "ellalgo_rs_b200", "csrc")
BUILD = os.path.join(CPP, "_build")
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
SRC = os.path.join(CPP, "emu_ellcta.cpp")
SRC_ST = os.path.join(CPP, "emu_stable_large.cpp")
SRC_BW = os.path.join(CPP, "emu_batch_warp.cpp")
SRC_LMI = os.path.join(CPP, "emu_lmi.cpp")
LIB_LMI, TSAN_LMI = os.path.join(BUILD, "libemu_lmi.so"), os.path.join(BUILD, "emu_lmi_tsan")
LIB_BW, TSAN_BW = os.path.join(BUILD, "libemu_batch_warp.so"), os.path.join(BUILD, "emu_batch_warp_tsan")
LIB, TSAN = os.path.join(BUILD, "libemu_ellcta.so"), os.path.join(BUILD, "emu_ellcta_tsan")
LIB_ST, TSAN_ST = os.path.join(BUILD, "libemu_stable_large.so"), os.path.join(BUILD, "emu_stable_large_tsan")
DEPS = [os.path.join(CPP, "cuda_emu.h")] + [os.path.join(CSRC, f) for f in ("ellcta_device.cuh", "stable_large_device.cuh", "batch_warp.cu",
                                                                              "batch_stable.cu", "lmi_device.cuh", "common.cuh",
                                                                              "ell_calc.cuh")]


def _build(out, extra, src=SRC):
    os.makedirs(BUILD, exist_ok=True)
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in DEPS + [src]):
        subprocess.run(["g++", "-std=c++17", "-g", "-ffp-contract=off", "-Wno-attributes", "-Wno-unknown-pragmas", "-I", CUDA_INC, src] + extra +
                       ["-o", out, "-lpthread", "-ldl"], check=True, capture_output=True, text=True)
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
    return _build(TSAN, ["-O1", "-fsanitize=thread", "-DEMU_MAIN", "-DEMU_JITTER"])


def _races_with_dropped_barrier(exe, k, var="EMU_DROP_BARRIER"):
    """every emulated thread skips its k-th __syncthreads (or __syncwarp); returns whether ThreadSanitizer reported a data race"""
    try:
        p = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=dict(os.environ, **{var: str(k)}))
        err = p.stderr
    except subprocess.TimeoutExpired as e:            # without the barrier the threads may also diverge and hang
        err = (e.stderr or b"").decode() if isinstance(e.stderr, bytes) else (e.stderr or "")
    return "WARNING: ThreadSanitizer: data race" in err


def test_no_data_race_under_thread_sanitizer():
    p = subprocess.run([_tsan_exe()], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "ThreadSanitizer" not in p.stderr, p.stdout[-2000:] + p.stderr[-4000:]
    assert "lowpass n=40: rc 0" in p.stdout and "3 of 3 with xbest" in p.stdout     # the shrink path ran too


def test_race_detector_sees_a_dropped_barrier():
    """sensitivity of the detector: barriers 1 … 13 of the bowl path (work queue, init, loop top, the oracle's two, update_core's
    four) were each dropped once by hand and each drop was reported; two of them are re-checked here"""
    assert _races_with_dropped_barrier(_tsan_exe(), 3)       # the oracle's barrier before the ordered sum of f
    assert _races_with_dropped_barrier(_tsan_exe(), 10)      # update_core: ratio*grad_t published before the rank-1


# ------------------------------------------------------------------------------------------ EllStable handle, 64 < n <= 2048
def _pack_upper(mq):
    n = mq.shape[0]
    return np.ascontiguousarray(np.concatenate([mq[r, r:] for r in range(n)]))


@pytest.mark.parametrize("n,start", [(70, "upper"), (100, "diag"), (65, "scalar")])
def test_emulated_ellstable_update_equals_checker(ref, n, start):  # src/ell_stable.rs:52-125 through stable_large_body
    from ellalgo_rs_b200 import _abi as A
    lib = C.CDLL(_build(LIB_ST, ["-O2", "-fPIC", "-shared"], SRC_ST))
    lib.emu_stable_large_update.restype = C.c_int
    lib.emu_stable_large_update.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double]
    rng = np.random.default_rng(n)
    xc = rng.uniform(-1, 1, n)
    kappa, mq, diag = 3.0, None, None
    if start == "diag":
        kappa, diag = 1.0, rng.uniform(0.5, 2.0, n)
    elif start == "upper":
        mq = np.eye(n) + np.triu(rng.uniform(-1e-3, 1e-3, (n, n)), 1)
    s = ref.Space(A.KIND_ELLSTABLE, n, kappa=kappa, mq=mq, diag=diag, xc=xc)
    qu = _pack_upper(mq if mq is not None else np.diag(diag if diag is not None else np.ones(n)))
    x, scal, status = xc.copy(), np.array([kappa, 0.0]), np.zeros(1, np.int32)
    cuts = [(A.UPDATE_BIAS, 0.01, 1, 0.10), (A.UPDATE_CENTRAL, 0.0, -1, 0.0), (A.UPDATE_BIAS, 1e6, -1, 0.0),     # the third fails: NoSoln
            (A.UPDATE_BIAS, 0.02, -1, 0.0), (A.UPDATE_CENTRAL, 0.0, 1, 0.05)]
    for k, (which, b0, has, b1) in enumerate(cuts):
        g = np.ascontiguousarray(rng.uniform(-1, 1, n) / np.sqrt(n))
        assert lib.emu_stable_large_update(qu.ctypes.data, x.ctypes.data, scal.ctypes.data, status.ctypes.data, g.ctypes.data, n, 1,
                                           which, b0, has, b1) == 0
        st = s.update(which, g, b0, b1 if has > 0 else None, single=has < 0)
        r = s.state()
        assert int(status[0]) == st, f"step {k}: status"
        assert scal[1] == r["tsq"] and scal[0] == r["kappa"] and np.array_equal(x, r["xc"]), f"step {k}"
    assert np.array_equal(qu, _pack_upper(s.state()["mq"]))      # the factor: diagonal and strict upper triangle


def test_ellstable_update_no_data_race_under_thread_sanitizer():
    # 128 emulated threads: ThreadSanitizer follows at most 256 running threads exactly (beyond that it misses races — tried)
    exe = _build(TSAN_ST, ["-O1", "-fsanitize=thread", "-DEMU_MAIN", "-DEMU_JITTER", "-DELL_STABLE_LARGE_THREADS=128"], SRC_ST)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "ThreadSanitizer" not in p.stderr, p.stdout[-2000:] + p.stderr[-4000:]
    assert "step 2: status 1" in p.stdout and "step 3: status 0" in p.stdout
    assert _races_with_dropped_barrier(exe, 5)            # forward substitution: w[j] final before column j is applied
    assert _races_with_dropped_barrier(exe, 150)          # back substitution: the row's terms before thread 0 sums them


# ------------------------------------------------------------------------------------------ warp-per-problem kernels, n <= 32
@pytest.fixture(scope="module")
def emu_warp():
    lib = C.CDLL(_build(LIB_BW, ["-O2", "-fPIC", "-shared"], SRC_BW))
    lib.emu_batch_warp_solve.restype = C.c_int
    lib.emu_batch_warp_solve.argtypes = [C.c_void_p, C.c_void_p]

    def solve(prob):
        out = BatchOutput.empty(prob.count, prob.n)
        d, r = prob.to_desc(0), out.to_result()
        assert lib.emu_batch_warp_solve(C.byref(d), C.byref(r)) == 0
        return out
    return solve


def test_emulated_warp_kernels_equal_checker(ref, emu_warp):  # csrc/batch_warp.cu compiled for the host
    rng = np.random.default_rng(11)
    for n, count, iters in ((7, 5, 200), (20, 5, 150), (32, 6, 150)):          # generic and FIXED32 variants
        prob = W.bowl(count, n=n, max_iters=iters)
        assert_parity(emu_warp(prob), ref.solve(prob), exact=True, what=f"emulated warp bowl n={n}")
        prob.kappa0 = np.ones(1)
        prob.mq0 = np.ascontiguousarray(100.0 * (np.eye(n)[None] + rng.uniform(-1, 1, (count, n, n)) / n))   # non-symmetric start
        prob.no_defer_trick = True
        assert_parity(emu_warp(prob), ref.solve(prob), exact=True, what=f"emulated warp bowl n={n} matrix start")
    for n, count, iters in ((12, 3, 400), (31, 2, 300)):
        prob = W.lowpass(count, ref.lowpass_table(n), n=n, max_iters=iters)
        assert_parity(emu_warp(prob), ref.solve(prob), exact=True, what=f"emulated warp lowpass n={n}")
    n = 32                                               # n = 32 from a matrix start: the variant the TMEM kernel leaves to this file
    prob = W.lowpass(2, ref.lowpass_table(n), n=n, max_iters=500)
    a = rng.uniform(-1, 1, (n, n)) / (4 * n)
    prob.mq0 = np.ascontiguousarray(np.eye(n) + a + a.T)
    assert_parity(emu_warp(prob), ref.solve(prob), exact=True, what="emulated warp lowpass n=32 matrix start")


def test_warp_kernels_no_data_race_under_thread_sanitizer():
    """lanes exchange vectors through shared memory ordered only by __syncwarp (a per-warp barrier here): none may be missing"""
    exe = _build(TSAN_BW, ["-O1", "-fsanitize=thread", "-DEMU_MAIN", "-DEMU_JITTER"], SRC_BW)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "ThreadSanitizer" not in p.stderr, p.stdout[-2000:] + p.stderr[-4000:]
    assert "lowpass n=32: rc 0" in p.stdout and "3 of 3 with xbest" in p.stdout
    # sensitivity: each of the first twelve __syncwarp calls was dropped once by hand, every drop was reported; two re-checked here
    assert _races_with_dropped_barrier(exe, 1, "EMU_DROP_SYNCWARP")       # the oracle's, before the ordered sum of f
    assert _races_with_dropped_barrier(exe, 4, "EMU_DROP_SYNCWARP")       # update_core's last, before the rank-1 reads ratio*grad_t


# ------------------------------------------------------------------------------------------ EllStable + LMI (configs[3])
def test_emulated_lmi_kernel_equals_checker(ref):  # csrc/batch_stable.cu + lmi_device.cuh compiled for the host
    from ellalgo_rs_b200 import _abi as A
    lib = C.CDLL(_build(LIB_LMI, ["-O2", "-fPIC", "-shared"], SRC_LMI))
    lib.emu_lmi_solve.restype = C.c_int
    lib.emu_lmi_solve.argtypes = [C.c_void_p, C.c_void_p]

    def solve(prob):
        out = BatchOutput.empty(prob.count, prob.n)
        d, r = prob.to_desc(0), out.to_result()
        assert lib.emu_lmi_solve(C.byref(d), C.byref(r)) == 0
        return out
    for n, m, count in ((64, 16, 3), (20, 7, 4), (33, 17, 2)):          # m <= 16: two LDL^T rows per step; m = 17: one row per step
        prob = W.lmi(count, n=n, m=m)
        prob.max_iters = 150
        g, r = solve(prob), ref.solve(prob)
        assert_parity(g, r, exact=True, what=f"emulated LMI n={n} m={m}")
        prob.oracle_id = A.ORACLE_LMI_OLD
        assert_parity(solve(prob), ref.solve(prob), exact=True, what=f"emulated LMIOld n={n} m={m}")
    assert g.num_iters.max() > 3
    prob = W.lmi(3, n=40, m=9, general=True, upper_scale=1e-6)          # EllStable::new_with_matrix: the GENERAL factor variant
    prob.max_iters = 60
    assert_parity(solve(prob), ref.solve(prob), exact=True, what="emulated LMI general start")
    prob = W.lmi0(4, n=24, m=8)
    prob.max_iters = 200
    g, r = solve(prob), ref.solve(prob)
    assert_parity(g, r, exact=True, what="emulated LMI0")
    assert (g.exit_status == A.EXIT_FEASIBLE).any()


def test_lmi_kernel_no_data_race_under_thread_sanitizer():
    """24 __syncwarp per iteration order the factor rows, the witness and the EllStable vectors between lanes"""
    exe = _build(TSAN_LMI, ["-O1", "-fsanitize=thread", "-DEMU_MAIN", "-DEMU_JITTER"], SRC_LMI)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "ThreadSanitizer" not in p.stderr, p.stdout[-2000:] + p.stderr[-4000:]
    assert "LMI0 diagonal start: rc 0" in p.stdout and "4 of 4 feasible" in p.stdout
    # sensitivity: the first 48 __syncwarp calls were each dropped once by hand and 47 drops were reported (the 48th, the pair
    # loop's closing __syncwarp in its first round, is covered by the next step's); two re-checked here
    assert _races_with_dropped_barrier(exe, 2, "EMU_DROP_SYNCWARP")
    assert _races_with_dropped_barrier(exe, 22, "EMU_DROP_SYNCWARP")      # the two-rows-per-step LDL^T: end of a step
