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
SRC_LP = os.path.join(CPP, "emu_lowpass32.cpp")
SRC_E64 = os.path.join(CPP, "emu_ell64.cpp")
SRC_SM = os.path.join(CPP, "emu_batch_small.cpp")
SRC_STR = os.path.join(CPP, "emu_stream.cpp")
SRC_SYM = os.path.join(CPP, "emu_sym.cpp")
LIB_SYM, TSAN_SYM = os.path.join(BUILD, "libemu_sym.so"), os.path.join(BUILD, "emu_sym_tsan")
LIB_STR, TSAN_STR = os.path.join(BUILD, "libemu_stream.so"), os.path.join(BUILD, "emu_stream_tsan")
LIB_SM = os.path.join(BUILD, "libemu_batch_small.so")
LIB_E64, TSAN_E64 = os.path.join(BUILD, "libemu_ell64.so"), os.path.join(BUILD, "emu_ell64_tsan")
LIB_LP, TSAN_LP = os.path.join(BUILD, "libemu_lowpass32.so"), os.path.join(BUILD, "emu_lowpass32_tsan")
LIB_LMI, TSAN_LMI = os.path.join(BUILD, "libemu_lmi.so"), os.path.join(BUILD, "emu_lmi_tsan")
LIB_BW, TSAN_BW = os.path.join(BUILD, "libemu_batch_warp.so"), os.path.join(BUILD, "emu_batch_warp_tsan")
LIB, TSAN = os.path.join(BUILD, "libemu_ellcta.so"), os.path.join(BUILD, "emu_ellcta_tsan")
LIB_ST, TSAN_ST = os.path.join(BUILD, "libemu_stable_large.so"), os.path.join(BUILD, "emu_stable_large_tsan")
DEPS = [os.path.join(CPP, "cuda_emu.h"), os.path.join(CPP, "tmem_emu.h"), os.path.join(CSRC, "batch_lowpass.cu"),
        os.path.join(CSRC, "batch_ell64.cu"), os.path.join(CSRC, "batch_small.cu"), os.path.join(CSRC, "glibc_math.cuh"),
        os.path.join(CPP, "stream_ptx_emu.h"), os.path.join(CSRC, "stream_update.cu"), os.path.join(CSRC, "stream_common.cuh"),
        os.path.join(CSRC, "stream_sym.cu")] + [os.path.join(CSRC, f) for f in ("ellcta_device.cuh", "stable_large_device.cuh", "batch_warp.cu",
                                                                              "batch_stable.cu", "lmi_device.cuh", "common.cuh",
                                                                              "ell_calc.cuh")]


_TSAN_OK = None


def _have_tsan():
    """g++ -fsanitize=thread can build and run a trivial program here (the runtime is an optional package)"""
    global _TSAN_OK
    if _TSAN_OK is None:
        os.makedirs(BUILD, exist_ok=True)
        probe = os.path.join(BUILD, "tsan_probe")
        with open(probe + ".cpp", "w") as f:
            f.write("int main() { return 0; }\n")
        try:
            subprocess.run(["g++", "-fsanitize=thread", probe + ".cpp", "-o", probe], check=True, capture_output=True)
            _TSAN_OK = subprocess.run([probe], capture_output=True, timeout=60).returncode == 0
        except Exception:
            _TSAN_OK = False
    return _TSAN_OK


def _build(out, extra, src=SRC):
    os.makedirs(BUILD, exist_ok=True)
    if "-fsanitize=thread" in extra and not _have_tsan():
        pytest.skip("no ThreadSanitizer runtime for g++ on this box")
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


def _races_with_dropped_barrier(exe, k, var="EMU_DROP_BARRIER", attempts=3):
    """every emulated thread skips its k-th __syncthreads (or __syncwarp); returns whether ThreadSanitizer reported a data race.
    Detection depends on the two unordered accesses really overlapping in some schedule, so a miss is retried."""
    for _ in range(attempts):
        try:
            p = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=dict(os.environ, **{var: str(k)}))
            err = p.stderr
        except subprocess.TimeoutExpired as e:            # without the barrier the threads may also diverge and hang
            err = (e.stderr or b"").decode() if isinstance(e.stderr, bytes) else (e.stderr or "")
        if "WARNING: ThreadSanitizer: data race" in err:
            return True
    return False


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


@pytest.mark.parametrize("n,start", [(70, "upper"), (100, "diag"), (65, "scalar"), (7, "diag"), (33, "upper"), (64, "upper")])
def test_emulated_ellstable_update_equals_checker(ref, n, start):  # src/ell_stable.rs:52-125: one CTA (n > 64) / one warp (n <= 64)
    from ellalgo_rs_b200 import _abi as A
    if n > 64:
        update = C.CDLL(_build(LIB_ST, ["-O2", "-fPIC", "-shared"], SRC_ST)).emu_stable_large_update
    else:
        update = C.CDLL(_build(LIB_LMI, ["-O2", "-fPIC", "-shared"], SRC_LMI)).emu_stable_single_update
    update.restype = C.c_int
    update.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double]
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
        assert update(qu.ctypes.data, x.ctypes.data, scal.ctypes.data, status.ctypes.data, g.ctypes.data, n, 1, which, b0, has, b1) == 0
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
    assert "single handle step 1: status 1" in p.stdout and "single handle step 2: status 0" in p.stdout
    # sensitivity: the first 48 __syncwarp calls were each dropped once by hand and 47 drops were reported (the 48th, the pair
    # loop's closing __syncwarp in its first round, is covered by the next step's); two re-checked here
    assert _races_with_dropped_barrier(exe, 2, "EMU_DROP_SYNCWARP")
    assert _races_with_dropped_barrier(exe, 22, "EMU_DROP_SYNCWARP")      # the two-rows-per-step LDL^T: end of a step


# ------------------------------------------------------------------------------------------ the headline kernel (configs[2])
def test_emulated_headline_kernel_equals_checker_and_golden(ref):  # csrc/batch_lowpass.cu with tensor memory emulated (tmem_emu.h)
    import golden_cases
    lib = C.CDLL(_build(LIB_LP, ["-O2", "-fPIC", "-shared"], SRC_LP))
    lib.emu_lowpass32_solve.restype = C.c_int
    lib.emu_lowpass32_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int]

    def solve(prob, variant=0):
        out = BatchOutput.empty(prob.count, prob.n)
        d, r = prob.to_desc(0), out.to_result()
        assert lib.emu_lowpass32_solve(C.byref(d), C.byref(r), variant) == 0
        return out
    t = ref.lowpass_table(32)
    prob = W.lowpass(13, t, max_iters=450)                 # more problems than emulated warps: the work queue's second round
    assert_parity(solve(prob), ref.solve(prob), exact=True, what="emulated lowpass32")
    few = W.lowpass(6, t, max_iters=450)
    want = ref.solve(few)
    for variant in (1, 2):                                 # ELLCUDA_LOWPASS_VARIANT: the spectrum's tail columns in TMEM
        assert_parity(solve(few, variant), want, exact=True, what=f"emulated lowpass32 variant {variant}")
    prob = few
    prob.use_parallel_cut = False
    prob.no_defer_trick = True
    assert_parity(solve(prob), ref.solve(prob), exact=True, what="emulated lowpass32 flags")
    bands = golden_cases.CASES["lowpass32_bands"](ref)     # short bands, empty transition band, edges on multiples of 32
    g = solve(bands)
    assert_parity(g, ref.solve(bands), exact=True, what="emulated lowpass32 band layouts")
    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
    for f in ("num_iters", "exit_status", "gamma", "xc", "xbest"):
        assert np.array_equal(getattr(g, f), gold[f"lowpass32_bands/{f}"], equal_nan=True), f


def test_headline_kernel_no_data_race_under_thread_sanitizer():
    """lanes share x, g, grad_t through shared memory (__syncwarp); warps share the SM's tensor memory (disjoint cells)"""
    exe = _build(TSAN_LP, ["-O1", "-fsanitize=thread", "-DEMU_MAIN", "-DEMU_JITTER"], SRC_LP)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "ThreadSanitizer" not in p.stderr, p.stdout[-2000:] + p.stderr[-4000:]
    assert p.stdout.count("9 of 9 with xbest") == 2
    # sensitivity: each of the first 20 __syncwarp calls and both start-up __syncthreads were dropped once by hand; all reported
    assert _races_with_dropped_barrier(exe, 3, "EMU_DROP_SYNCWARP")
    assert _races_with_dropped_barrier(exe, 0)               # the barrier that publishes the staged spectrum


# ------------------------------------------------------------------------------------------ Ell with M in tensor memory, n <= 64
def test_emulated_ell64_kernel_equals_checker(ref):  # csrc/batch_ell64.cu with tensor memory emulated
    from ellalgo_rs_b200 import _abi as A
    lib = C.CDLL(_build(LIB_E64, ["-O2", "-fPIC", "-shared"], SRC_E64))
    lib.emu_ell64_solve.restype = C.c_int
    lib.emu_ell64_solve.argtypes = [C.c_void_p, C.c_void_p]

    def solve(prob):
        out = BatchOutput.empty(prob.count, prob.n)
        d, r = prob.to_desc(0), out.to_result()
        assert lib.emu_ell64_solve(C.byref(d), C.byref(r)) == 0
        return out
    rng = np.random.default_rng(64)
    for n, count, iters in ((33, 7, 120), (48, 9, 100), (64, 8, 100)):
        prob = W.bowl(count, n=n, max_iters=iters)
        assert_parity(solve(prob), ref.solve(prob), exact=True, what=f"emulated ell64 bowl n={n}")
    n = 50
    prob = W.bowl(5, n=n, max_iters=80)
    prob.kappa0 = np.ones(1)
    prob.mq0 = np.ascontiguousarray(100.0 * (np.eye(n)[None] + rng.uniform(-1, 1, (5, n, n)) / n))     # non-symmetric start
    for nd in (False, True):
        prob.no_defer_trick = nd
        assert_parity(solve(prob), ref.solve(prob), exact=True, what=f"emulated ell64 bowl matrix start no_defer={nd}")
    prob = W.lowpass(3, ref.lowpass_table(40), n=40, max_iters=400)
    assert_parity(solve(prob), ref.solve(prob), exact=True, what="emulated ell64 lowpass n=40")
    for make, oracle in ((lambda: W.lmi(5, n=64, m=16), A.ORACLE_LMI), (lambda: W.lmi(4, n=20, m=7), A.ORACLE_LMI_OLD),
                         (lambda: W.lmi0(5, n=40, m=9), A.ORACLE_LMI0)):       # tests/lmi_tests.rs:174-225 run both spaces
        prob = make()
        prob.kind, prob.oracle_id, prob.max_iters = A.KIND_ELL, oracle, 150
        g, r = solve(prob), ref.solve(prob)
        assert_parity(g, r, exact=True, what=f"emulated ell64 LMI oracle {oracle} on Ell")
    assert (g.exit_status == A.EXIT_FEASIBLE).any()


def test_ell64_kernel_no_data_race_under_thread_sanitizer():
    exe = _build(TSAN_E64, ["-O1", "-fsanitize=thread", "-DEMU_MAIN", "-DEMU_JITTER"], SRC_E64)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "ThreadSanitizer" not in p.stderr, p.stdout[-2000:] + p.stderr[-4000:]
    assert "LMI0 on Ell n=40 m=9: rc 0" in p.stdout and "7 of 7 with xbest" in p.stdout
    # sensitivity: the first 16 __syncwarp calls were each dropped once by hand, 15 reported (the first only orders the zero fill
    # of the vectors, which the first iteration's own __syncwarp covers); one re-checked here
    assert _races_with_dropped_barrier(exe, 2, "EMU_DROP_SYNCWARP")


# ------------------------------------------------------------------------------------------ the n = 2 kernels: every pinned count
def test_emulated_n2_kernels_reproduce_every_pinned_count(ref):  # csrc/batch_small.cu through its own dispatch table
    import math
    from ellalgo_rs_b200 import _abi as A
    from ellalgo_rs_b200.problem import BatchProblem
    lib = C.CDLL(_build(LIB_SM, ["-O2", "-fPIC", "-shared", "-mfma"], SRC_SM))    # -mfma: glibc's exp/log use a true fma
    lib.emu_batch_small_solve.restype = C.c_int
    lib.emu_batch_small_solve.argtypes = [C.c_void_p, C.c_void_p]

    def solve(prob):
        out = BatchOutput.empty(prob.count, prob.n)
        d, r = prob.to_desc(0), out.to_result()
        assert lib.emu_batch_small_solve(C.byref(d), C.byref(r)) == 0
        return out

    def check(prob, what, pinned=None):
        g = solve(prob)
        assert_parity(g, ref.solve(prob), exact=True, what=what)
        if pinned is not None:
            assert g.num_iters[0] == pinned, f"{what}: {g.num_iters[0]} iterations, the reference's test pins {pinned}"
        return g
    inf = math.inf
    check(W.example1(), "example1", 25)                                                            # src/example1.rs:39-50
    check(BatchProblem(oracle_id=A.ORACLE_EXAMPLE1_RR, n=2, count=3, kappa0=10.0, gamma0=-inf, tolerance=1e-10), "example1_rr", 25)
    check(BatchProblem(oracle_id=A.ORACLE_EXAMPLE4, n=2, count=3, kappa0=10.0, gamma0=-inf, tolerance=1e-10), "example4", 82)
    g = check(W.profit(300), "profit", 83)                                                         # profit_oracle.rs:207-223
    assert g.xbest[0, 0] <= math.log(30.5)
    p = W.profit(1)
    p.oracle_id = A.ORACLE_PROFIT_RB
    p.params = np.concatenate([W.PROFIT_REF, [0.003, 0.007, 1.0, 1.0, 1.0]])
    check(p, "profit_rb", 90)                                                                      # :225-242
    p = W.profit(70)
    p.oracle_id, p.mode = A.ORACLE_PROFIT_Q, A.MODE_OPTIM_Q
    check(p, "profit_q", 29)                                                                       # :244-260
    check(BatchProblem(oracle_id=A.ORACLE_QUASICVX, n=2, count=1, kappa0=1.0, diag0=[10.0, 10.0], gamma0=0.0, tolerance=1e-8),
          "quasicvx", 35)                                                                          # src/quasicvx.rs:61-78
    check(BatchProblem(oracle_id=A.ORACLE_QUASICVX, n=2, count=1, kind=A.KIND_ELLSTABLE, kappa0=1.0, diag0=[10.0, 10.0], gamma0=0.0,
                       tolerance=1e-8), "quasicvx on EllStable", 49)
    g = check(BatchProblem(oracle_id=A.ORACLE_EXAMPLE3, n=2, count=1, mode=A.MODE_BSEARCH, kappa0=100.0, gamma0=-100.0,
                           params=[100.0, 1e-8, 2000.0], tolerance=1e-8), "bsearch example3", 34)  # src/example3.rs:68-85
    assert g.has_xbest[0] == 1
    for kind in (A.KIND_ELL, A.KIND_ELLSTABLE):                                                    # tests/cutting_plane_tests.rs, example2
        for what, kw, pinned in (("always_cut_feas", dict(oracle_id=A.ORACLE_ALWAYS_CUT, mode=A.MODE_FEAS, kappa0=10.0, max_iters=200), 2),
                                 ("always_cut_q", dict(oracle_id=A.ORACLE_ALWAYS_CUT, mode=A.MODE_OPTIM_Q, kappa0=10.0, gamma0=0.0, max_iters=5), 2),
                                 ("halfplane", dict(oracle_id=A.ORACLE_HALFPLANE, mode=A.MODE_FEAS, kappa0=10.0, max_iters=200), 0),
                                 ("optim_box", dict(oracle_id=A.ORACLE_OPTIM_BOX, kappa0=10.0, gamma0=0.0, max_iters=200), None),
                                 ("optim_box_q", dict(oracle_id=A.ORACLE_OPTIM_BOX, mode=A.MODE_OPTIM_Q, kappa0=10.0, gamma0=0.0, max_iters=200), None),
                                 ("example2", dict(oracle_id=A.ORACLE_EXAMPLE2, mode=A.MODE_FEAS, kappa0=10.0), 1),
                                 ("example2_inf", dict(oracle_id=A.ORACLE_EXAMPLE2, mode=A.MODE_FEAS, kappa0=10.0, xc0=[100.0, 100.0]), 0)):
            check(BatchProblem(n=2, count=3, kind=kind, **kw), f"{what} kind={kind}", pinned if kind == A.KIND_ELL else None)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))                         # and the committed vectors
    import golden_cases
    for name in ("example1", "profit", "profit_q", "quasicvx_stable"):
        g = solve(golden_cases.CASES[name](ref))
        for f in ("num_iters", "exit_status", "gamma", "xc", "xbest"):
            assert np.array_equal(getattr(g, f), gold[f"{name}/{f}"], equal_nan=True), f"{name}/{f}"


# ------------------------------------------------------------------------------------------ the large-n path (configs[4])
class _EmuEll:
    """Ell::update_core through the emulated sweep_kernel + finish_kernel (tests/cpp/emu_stream.cpp)"""

    def __init__(self, lib, n, kappa, xc, mq=None, diag=None, no_defer=False, use_parallel=True, reduce_mode=0, cols=0, rows=64,
                 warps=1, ll=False):
        self.lib, self.n = lib, n
        mq = None if mq is None else np.ascontiguousarray(mq, dtype=np.float64)
        diag = None if diag is None else np.ascontiguousarray(diag, dtype=np.float64)
        xc = np.ascontiguousarray(xc, dtype=np.float64)
        self.h = lib.emu_ell_new(n, kappa, None if mq is None else mq.ctypes.data, None if diag is None else diag.ctypes.data,
                                 xc.ctypes.data, int(no_defer), int(use_parallel), reduce_mode, cols, rows, warps, int(ll))

    def update(self, which, g, b0, b1=None, single=True):
        g = np.ascontiguousarray(g, dtype=np.float64)
        return self.lib.emu_ell_update(self.h, which, g.ctypes.data, float(b0), 1 if b1 is not None else (-1 if single else 0),
                                       0.0 if b1 is None else float(b1))

    def state(self):
        xc, mq = np.zeros(self.n), np.zeros((self.n, self.n))
        k, t = C.c_double(), C.c_double()
        self.lib.emu_ell_state(self.h, xc.ctypes.data, mq.ctypes.data, C.byref(k), C.byref(t))
        return dict(xc=xc, mq=mq, kappa=k.value, tsq=t.value)

    def dot_mv(self, x):
        x, y = np.ascontiguousarray(x, dtype=np.float64), np.zeros(self.n)
        self.lib.emu_ell_dot_mv(self.h, x.ctypes.data, y.ctypes.data)
        return y

    def __del__(self):
        self.lib.emu_ell_free(self.h)


@pytest.fixture(scope="module")
def emu_stream():
    lib = C.CDLL(_build(LIB_STR, ["-O2", "-fPIC", "-shared"], SRC_STR))
    lib.emu_ell_new.restype = C.c_void_p
    lib.emu_ell_new.argtypes = [C.c_longlong, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 7
    lib.emu_ell_update.restype = C.c_int
    lib.emu_ell_update.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_double]
    lib.emu_ell_state.argtypes = [C.c_void_p] * 5
    lib.emu_ell_dot_mv.argtypes = [C.c_void_p] * 3
    lib.emu_ell_free.argtypes = [C.c_void_p]
    lib.emu_group_new.restype = C.c_void_p
    lib.emu_group_new.argtypes = [C.c_int, C.c_longlong, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 4
    lib.emu_group_update.restype = C.c_int
    lib.emu_group_update.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_double]
    lib.emu_group_state.restype = C.c_int
    lib.emu_group_state.argtypes = [C.c_void_p] * 5
    lib.emu_group_free.argtypes = [C.c_void_p]
    return lib


LARGE_N_VARIANTS = [  # n, reduce_mode, cols, tile rows, warps, flagged lines, no_defer, start
    (150, 0, 64, 64, 1, False, False, "scalar"), (129, 0, 64, 64, 2, True, False, "diag"), (100, 0, 32, 64, 1, True, True, "dense"),
    (90, 0, 16, 64, 1, False, False, "dense"), (140, 0, 64, 32, 1, False, True, "scalar"), (64, 0, 0, 64, 1, False, False, "dense"),
    (200, 3, 64, 64, 1, True, False, "diag"), (130, 2, 32, 64, 1, False, True, "dense"), (77, 0, 0, 64, 1, True, False, "scalar")]


@pytest.mark.parametrize("n,reduce_mode,cols,rows,warps,ll,no_defer,start", LARGE_N_VARIANTS)
def test_emulated_large_n_update_equals_checker(ref, emu_stream, n, reduce_mode, cols, rows, warps, ll, no_defer, start):
    """src/ell.rs:97-137 through sweep_kernel + finish_kernel: bit for bit in reduce_mode 0, 1e-9 with row segments"""
    from ellalgo_rs_b200 import _abi as A
    rng = np.random.default_rng(n + reduce_mode)
    xc = rng.uniform(-1, 1, n)
    kappa, mq, diag = 2.5, None, None
    if start == "diag":
        kappa, diag = 1.0, rng.uniform(0.5, 2.0, n)
    elif start == "dense":
        a = rng.uniform(-1, 1, (n, n)) / n
        kappa, mq = 1.5, np.eye(n) + a + a.T
    e = _EmuEll(emu_stream, n, kappa, xc, mq=mq, diag=diag, no_defer=no_defer, reduce_mode=reduce_mode, cols=cols, rows=rows, warps=warps, ll=ll)
    s = ref.Space(A.KIND_ELL, n, kappa=kappa, mq=mq, diag=diag, xc=xc, no_defer_trick=no_defer)
    exact = reduce_mode == 0
    cuts = [(A.UPDATE_BIAS, 0.01, 0.10, False), (A.UPDATE_CENTRAL, 0.0, None, True), (A.UPDATE_BIAS, 1e6, None, True),      # 3rd: NoSoln
            (A.UPDATE_BIAS, 0.02, None, True), (A.UPDATE_Q, -0.5, None, False), (A.UPDATE_CENTRAL, 0.0, 0.05, False)]
    for k, (which, b0, b1, single) in enumerate(cuts):
        g = rng.uniform(-1, 1, n) / np.sqrt(n)
        assert e.update(which, g, b0, b1, single) == s.update(which, g, b0, b1, single), f"step {k}: status"
    got, want = e.state(), s.state()
    for f in ("tsq", "kappa", "xc", "mq"):
        if exact:
            assert np.array_equal(got[f], want[f]), f"{f} not bit-identical"
        else:
            assert np.allclose(got[f], want[f], rtol=1e-9, atol=1e-12), f
    if exact:                                             # Arr::dot_mv on the handle (power iteration / CG use it): column sums of M
        x = rng.uniform(-1, 1, n)
        m = want["mq"]
        y = np.zeros(n)
        for j in range(n):                                # out[i] = sum_j M[i][j] x[j], j ascending from 0.0 (src/arr.rs:426-442)
            y = y + m[:, j] * x[j]
        assert np.array_equal(e.dot_mv(x), y)


@pytest.mark.parametrize("nranks,n,reduce_mode,cols,no_defer", [(2, 130, 0, 64, False), (4, 200, 0, 0, True), (3, 150, 3, 64, False),
                                                                  (8, 192, 0, 16, False), (2, 100, 2, 32, True)])
def test_emulated_row_sharded_group_equals_checker(ref, emu_stream, nranks, n, reduce_mode, cols, no_defer):
    """the multi-GPU form of the large-n path with the real kernels and no GPU: rank r keeps rows [r n/P, (r+1) n/P) of
    M, every sweep stores its slab of grad_t as flagged lines into every rank's exchange block (epoch-parity double buffer),
    every finish polls its own block; all ranks must end up with identical xc / kappa / tsq — and with the checker's"""
    from ellalgo_rs_b200 import _abi as A
    rng = np.random.default_rng(10 * n + nranks)
    xc = rng.uniform(-1, 1, n)
    a = rng.uniform(-1, 1, (n, n)) / n
    mq = np.ascontiguousarray(np.eye(n) + a + a.T)
    g = emu_stream.emu_group_new(nranks, n, 1.5, mq.ctypes.data, None, np.ascontiguousarray(xc).ctypes.data, int(no_defer), 1, reduce_mode, cols)
    s = ref.Space(A.KIND_ELL, n, kappa=1.5, mq=mq, xc=xc, no_defer_trick=no_defer)
    try:
        cuts = [(A.UPDATE_BIAS, 0.01, 0.10, False), (A.UPDATE_CENTRAL, 0.0, None, True), (A.UPDATE_BIAS, 1e6, None, True),
                (A.UPDATE_BIAS, 0.02, None, True), (A.UPDATE_CENTRAL, 0.0, 0.05, False)]
        for k, (which, b0, b1, single) in enumerate(cuts):
            grad = np.ascontiguousarray(rng.uniform(-1, 1, n) / np.sqrt(n))
            st = emu_stream.emu_group_update(g, which, grad.ctypes.data, float(b0), 1 if b1 is not None else (-1 if single else 0),
                                             0.0 if b1 is None else float(b1))
            assert st == s.update(which, grad, b0, b1, single), f"step {k}: status (or the ranks disagree)"
        gx, gm = np.zeros(n), np.zeros((n, n))
        gk, gt = C.c_double(), C.c_double()
        assert emu_stream.emu_group_state(g, gx.ctypes.data, gm.ctypes.data, C.byref(gk), C.byref(gt)) == 0, "ranks hold different xc / kappa / tsq"
        want = s.state()
        if reduce_mode == 0:
            assert gt.value == want["tsq"] and gk.value == want["kappa"]
            assert np.array_equal(gx, want["xc"]) and np.array_equal(gm, want["mq"])
        else:
            assert np.allclose([gt.value, gk.value], [want["tsq"], want["kappa"]], rtol=1e-9)
            assert np.allclose(gx, want["xc"], rtol=1e-9, atol=1e-12) and np.allclose(gm, want["mq"], rtol=1e-9, atol=1e-12)
    finally:
        emu_stream.emu_group_free(g)


def test_large_n_kernels_no_data_race_under_thread_sanitizer():
    """the mbarrier hand-off of the TMA ring, the CTA-wide staging of finish's ordered sum, the flagged lines"""
    exe = _build(TSAN_STR, ["-O1", "-fsanitize=thread", "-DEMU_MAIN", "-DEMU_JITTER", "-DELLCUDA_FIN_THREADS=128"], SRC_STR)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "ThreadSanitizer" not in p.stderr, p.stdout[-2000:] + p.stderr[-4000:]
    assert p.stdout.count("step 2: status 1") == 5 and p.stdout.count("step 3: status 0") == 5
    # sensitivity: the per-tile CTA barrier of the sweep (one warp: __syncwarp, two warps: __syncthreads) was dropped at several
    # positions and every drop was reported; the barrier after the mbarrier initialisation and the warp's re-convergence at the
    # end of a tile order no plain memory (no report)
    assert _races_with_dropped_barrier(exe, 1, "EMU_DROP_SYNCWARP")
    assert _races_with_dropped_barrier(exe, 1)


# ------------------------------------------------------------------------------------------ large n, symmetric storage
@pytest.mark.parametrize("nranks,n,start,use_map,no_defer,reduce_mode", [(1, 130, "scalar", 1, False, 0), (1, 200, "dense", 1, True, 0),
                                                                          (1, 100, "diag", 0, False, 2), (2, 200, "dense", 1, False, 0),
                                                                          (3, 260, "diag", 1, True, 0), (4, 330, "dense", 0, False, 0)])
def test_emulated_symmetric_storage_equals_checker(ref, nranks, n, start, use_map, no_defer, reduce_mode):
    """ELL_STORAGE_SYMMETRIC through sym_sweep / sym_combine (one handle) or sym_combine_push / sym_reduce (sharded by
    shard_rank / shard_count) + finish_kernel: within 1e-9 of the checker — the bar the device path is held to — and identical
    on every rank of a group"""
    from ellalgo_rs_b200 import _abi as A
    lib = C.CDLL(_build(LIB_SYM, ["-O2", "-fPIC", "-shared"], SRC_SYM))
    lib.emu_sym_new.restype = C.c_void_p
    lib.emu_sym_new.argtypes = [C.c_int, C.c_longlong, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 4
    lib.emu_sym_update.restype = C.c_int
    lib.emu_sym_update.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_double]
    lib.emu_sym_state.restype = C.c_int
    lib.emu_sym_state.argtypes = [C.c_void_p] * 5
    lib.emu_sym_free.argtypes = [C.c_void_p]
    rng = np.random.default_rng(7 * n + nranks)
    xc = np.ascontiguousarray(rng.uniform(-1, 1, n))
    kappa, mq, diag = 2.5, None, None
    if start == "diag":
        kappa, diag = 1.0, np.ascontiguousarray(rng.uniform(0.5, 2.0, n))
    elif start == "dense":
        a = rng.uniform(-1, 1, (n, n)) / n
        kappa, mq = 1.5, np.ascontiguousarray(np.eye(n) + a + a.T)
    g = lib.emu_sym_new(nranks, n, kappa, None if mq is None else mq.ctypes.data, None if diag is None else diag.ctypes.data, xc.ctypes.data,
                        int(no_defer), 1, reduce_mode, use_map)
    s = ref.Space(A.KIND_ELL, n, kappa=kappa, mq=mq, diag=diag, xc=xc, no_defer_trick=no_defer)
    try:
        cuts = [(A.UPDATE_BIAS, 0.01, 0.10, False), (A.UPDATE_CENTRAL, 0.0, None, True), (A.UPDATE_BIAS, 1e6, None, True),
                (A.UPDATE_BIAS, 0.02, None, True), (A.UPDATE_CENTRAL, 0.0, 0.05, False)]
        for k, (which, b0, b1, single) in enumerate(cuts):
            grad = np.ascontiguousarray(rng.uniform(-1, 1, n) / np.sqrt(n))
            st = lib.emu_sym_update(g, which, grad.ctypes.data, float(b0), 1 if b1 is not None else (-1 if single else 0), 0.0 if b1 is None else float(b1))
            assert st == s.update(which, grad, b0, b1, single), f"step {k}: status (or the ranks disagree)"
        gx, gm = np.zeros(n), np.zeros((n, n))
        gk, gt = C.c_double(), C.c_double()
        assert lib.emu_sym_state(g, gx.ctypes.data, gm.ctypes.data, C.byref(gk), C.byref(gt)) == 0, "ranks hold different xc / kappa / tsq"
        want = s.state()
        assert np.allclose([gt.value, gk.value], [want["tsq"], want["kappa"]], rtol=1e-9)
        assert np.allclose(gx, want["xc"], rtol=1e-9, atol=1e-12) and np.allclose(gm, want["mq"], rtol=1e-9, atol=1e-12)
    finally:
        lib.emu_sym_free(g)


def test_symmetric_storage_kernels_no_data_race_under_thread_sanitizer():
    """two warps share a tile between the passes (CTA barriers), eight warps share the ordered combine's partial sums"""
    exe = _build(TSAN_SYM, ["-O1", "-fsanitize=thread", "-DEMU_MAIN", "-DEMU_JITTER", "-DELLCUDA_FIN_THREADS=128"], SRC_SYM)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "ThreadSanitizer" not in p.stderr, p.stdout[-2000:] + p.stderr[-4000:]
    assert p.stdout.count("step 1: status 1") == 3 and p.stdout.count("step 2: status 0") == 3 and p.stdout.count("ranks differ 0") == 3
    assert _races_with_dropped_barrier(exe, 3)       # sensitivity: a CTA barrier of the tile loop
