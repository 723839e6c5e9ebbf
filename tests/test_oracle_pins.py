"""Pins the CPU checker (oracle/) against every iteration count and known-answer vector the reference's own
tests hold for the ellipsoid path.  CPU only."""
import ctypes as C
import math

import numpy as np
import pytest

from ellalgo_rs_b200 import _abi as A
from ellalgo_rs_b200 import workloads as W
from ellalgo_rs_b200.problem import BatchProblem

APPROX = dict(rel=0, abs=1e-6)  # approx_eq::assert_approx_eq! default eps = 1e-6


def one(ref, **kw):
    return ref.solve(BatchProblem(count=1, **kw))


# ---------------------------------------------------------------- §4(b): pinned iteration counts
def test_example1_feasible(ref):  # src/example1.rs:39-50
    o = ref.solve(W.example1())
    assert o.num_iters[0] == 25 and o.has_xbest[0] == 1 and o.exit_status[0] == A.EXIT_TOL


def test_example1_infeasible(ref):  # src/example1.rs:52-70
    o = one(ref, oracle_id=A.ORACLE_EXAMPLE1, n=2, kappa0=1.0, diag0=[10.0, 10.0], xc0=[100.0, 100.0], gamma0=-math.inf)
    assert o.has_xbest[0] == 0
    o = one(ref, oracle_id=A.ORACLE_EXAMPLE1, n=2, kappa0=1.0, diag0=[10.0, 10.0], gamma0=100.0)
    assert o.has_xbest[0] == 0


def test_example1_rr(ref):  # src/example1_rr.rs:62-74
    o = one(ref, oracle_id=A.ORACLE_EXAMPLE1_RR, n=2, kappa0=10.0, gamma0=-math.inf, tolerance=1e-10)
    assert o.num_iters[0] == 25 and o.has_xbest[0] == 1


def test_example4(ref):  # src/example4.rs:65-77
    o = one(ref, oracle_id=A.ORACLE_EXAMPLE4, n=2, kappa0=10.0, gamma0=-math.inf, tolerance=1e-10)
    assert o.num_iters[0] == 82 and o.has_xbest[0] == 1


def test_example3_bsearch(ref):  # src/example3.rs:68-85
    f = C.c_int()
    it = ref.lib().ellref_bsearch_example3(100.0, np.zeros(2).ctypes.data, -100.0, 100.0, 1e-8, 2000, 1e-8, 2000, C.byref(f))
    assert (bool(f.value), it) == (True, 34)


def test_example3_bsearch_batched_form(ref):  # the same test through the batch descriptor's MODE_BSEARCH
    o = one(ref, oracle_id=A.ORACLE_EXAMPLE3, n=2, mode=A.MODE_BSEARCH, kappa0=100.0, gamma0=-100.0, params=[100.0, 1e-8, 2000.0],
            tolerance=1e-8)
    assert (o.has_xbest[0], o.num_iters[0], o.exit_status[0]) == (1, 34, A.EXIT_TOL)


def test_profit(ref):  # src/oracles/profit_oracle.rs:207-223
    o = ref.solve(W.profit(1))
    assert o.num_iters[0] == 83 and o.has_xbest[0] == 1 and o.xbest[0, 0] <= math.log(30.5)


def test_profit_rb(ref):  # :225-242
    p = W.profit(1)
    p.oracle_id = A.ORACLE_PROFIT_RB
    p.params = np.concatenate([W.PROFIT_REF, [0.003, 0.007, 1.0, 1.0, 1.0]])
    o = ref.solve(p)
    assert o.num_iters[0] == 90 and o.xbest[0, 0] <= math.log(30.5)


def test_profit_q(ref):  # :244-260
    p = W.profit(1)
    p.oracle_id, p.mode = A.ORACLE_PROFIT_Q, A.MODE_OPTIM_Q
    o = ref.solve(p)
    assert o.num_iters[0] == 29 and o.xbest[0, 0] <= math.log(30.5)


def test_profit_direct(ref):  # :262-277 (exact assert_eq!)
    g, beta, gamma = np.zeros(2), C.c_double(), C.c_double(0.0)
    p = W.PROFIT_REF.copy()
    y = np.array([3.5, 2.0])
    shrunk = ref.lib().ellref_profit_assess(p.ctypes.data, y.ctypes.data, C.byref(gamma), g.ctypes.data, C.byref(beta))
    assert shrunk == 0 and beta.value == 3.5 - math.log(30.5)
    y = np.array([3.0, 2.0])
    shrunk = ref.lib().ellref_profit_assess(p.ctypes.data, y.ctypes.data, C.byref(gamma), g.ctypes.data, C.byref(beta))
    assert shrunk == 1 and beta.value == 0.0


def test_quasicvx(ref):  # src/quasicvx.rs:61-78
    o = one(ref, oracle_id=A.ORACLE_QUASICVX, n=2, kappa0=1.0, diag0=[10.0, 10.0], gamma0=0.0, tolerance=1e-8)
    x = o.xbest[0]
    assert o.num_iters[0] == 35 and 0.49 <= x[0] * x[0] <= 0.51 and 1.6 <= math.exp(x[1]) <= 1.7


def test_quasicvx_stable_and_infeasible(ref):  # src/quasicvx.rs:80-131
    o = one(ref, oracle_id=A.ORACLE_QUASICVX, n=2, kind=A.KIND_ELLSTABLE, kappa0=1.0, diag0=[10.0, 10.0], gamma0=0.0, tolerance=1e-8)
    assert o.has_xbest[0] == 1
    for kind in (A.KIND_ELL, A.KIND_ELLSTABLE):
        o = one(ref, oracle_id=A.ORACLE_QUASICVX, n=2, kind=kind, kappa0=10.0, xc0=[100.0, 100.0], gamma0=0.0)
        assert o.has_xbest[0] == 0
        o = one(ref, oracle_id=A.ORACLE_QUASICVX, n=2, kind=kind, kappa0=1.0, diag0=[10.0, 10.0], gamma0=100.0)
        assert o.has_xbest[0] == 0


def test_cutting_plane_tests(ref):  # tests/cutting_plane_tests.rs:128-186, 266-327
    o = one(ref, oracle_id=A.ORACLE_HALFPLANE, n=2, mode=A.MODE_FEAS, kappa0=10.0, max_iters=200)
    assert (o.has_xbest[0], o.num_iters[0]) == (1, 0)
    for mi in (200, 5):
        o = one(ref, oracle_id=A.ORACLE_ALWAYS_CUT, n=2, mode=A.MODE_FEAS, kappa0=10.0, max_iters=mi)
        assert (o.has_xbest[0], o.num_iters[0]) == (0, 2)
    o = one(ref, oracle_id=A.ORACLE_OPTIM_BOX, n=2, kappa0=10.0, gamma0=0.0, max_iters=200)
    assert o.has_xbest[0] == 1
    o = one(ref, oracle_id=A.ORACLE_OPTIM_BOX, n=2, kappa0=10.0, gamma0=100.0, max_iters=4)
    assert (o.has_xbest[0], o.num_iters[0]) == (0, 0)
    o = one(ref, oracle_id=A.ORACLE_ALWAYS_CUT, n=2, kappa0=10.0, gamma0=0.0, max_iters=5)
    assert (o.has_xbest[0], o.num_iters[0]) == (0, 2)
    o = one(ref, oracle_id=A.ORACLE_OPTIM_BOX, n=2, mode=A.MODE_OPTIM_Q, kappa0=10.0, gamma0=0.0, max_iters=200)
    assert o.has_xbest[0] == 1
    o = one(ref, oracle_id=A.ORACLE_OPTIM_BOX, n=2, mode=A.MODE_OPTIM_Q, kappa0=10.0, gamma0=100.0, max_iters=20)
    assert (o.has_xbest[0], o.num_iters[0]) == (0, 0)
    o = one(ref, oracle_id=A.ORACLE_ALWAYS_CUT, n=2, mode=A.MODE_OPTIM_Q, kappa0=10.0, gamma0=0.0, max_iters=5)
    assert (o.has_xbest[0], o.num_iters[0]) == (0, 2)
    f = C.c_int()
    assert ref.lib().ellref_bsearch_bs2(-100.0, 100.0, 1e-7, 2000, C.byref(f)) == 30 and f.value == 1
    assert ref.lib().ellref_bsearch_bs2(-100.0, -50.0, 1e-20, 20, C.byref(f)) == 20 and f.value == 0


def test_example2(ref):  # tests/example2_tests.rs:50-68
    o = one(ref, oracle_id=A.ORACLE_EXAMPLE2, n=2, mode=A.MODE_FEAS, kappa0=10.0)
    assert (o.has_xbest[0], o.num_iters[0]) == (1, 1)
    o = one(ref, oracle_id=A.ORACLE_EXAMPLE2, n=2, mode=A.MODE_FEAS, kappa0=10.0, xc0=[100.0, 100.0])
    assert (o.has_xbest[0], o.num_iters[0]) == (0, 0)


LMI_TAB = np.array(
    [1.0, -1.0, 1.0]
    + [-7, -11, -11, 3, 7, -18, -18, 8, -2, -8, -8, 1] + [33, -9, -9, 26]
    + [-21, -11, 0, -11, 10, 8, 0, 8, 5, 0, 10, 16, 10, -10, -10, 16, -10, 3, -5, 2, -17, 2, -6, 8, -17, 8, 6]
    + [14, 9, 40, 9, 91, 10, 40, 10, 15], dtype=float)


def test_lmi_tests(ref):  # tests/lmi_tests.rs:174-225 (bounds <300 / <400; survey-derived exact values 11 / 21)
    o = one(ref, oracle_id=A.ORACLE_LMI_LINOPT, n=3, kappa0=10.0, gamma0=math.inf, table=LMI_TAB, aux=[2, 2, 3])
    assert o.has_xbest[0] == 1 and o.num_iters[0] < 300 and o.num_iters[0] == 11
    assert o.gamma[0] == -3.7502205782089257
    o = one(ref, oracle_id=A.ORACLE_LMI_LINOPT, n=3, kind=A.KIND_ELLSTABLE, kappa0=10.0, gamma0=math.inf, table=LMI_TAB, aux=[2, 2, 3])
    assert o.has_xbest[0] == 1 and o.num_iters[0] < 400 and o.num_iters[0] == 21
    assert o.gamma[0] == -3.5724046875091333
    # test_lmi_oracle (:60-64): x = 0 is feasible for (F1, B1)
    g, beta = np.zeros(3), C.c_double()
    f1, b1 = LMI_TAB[3:15].copy(), LMI_TAB[15:19].copy()
    assert ref.lib().ellref_lmi_assess(3, 2, f1.ctypes.data, b1.ctypes.data, np.zeros(3).ctypes.data, g.ctypes.data, C.byref(beta)) == 0


def _lmi_assess(ref, variant, f, b, x, m):
    f = np.ascontiguousarray(f, dtype=float)
    b = np.ascontiguousarray(b if b is not None else np.zeros(m * m), dtype=float)
    x = np.ascontiguousarray(x, dtype=float)
    g, beta = np.zeros(len(x)), C.c_double()
    cut = ref.lib().ellref_lmi_assess_variant(variant, len(x), m, f.ctypes.data, b.ctypes.data, x.ctypes.data, g.ctypes.data, C.byref(beta))
    return cut, g, beta.value


def test_lmi_old_and_lmi0_oracles(ref):
    """src/oracles/lmi_old_oracle.rs, lmi0_oracle.rs — tests/lmi_tests.rs:66-113 (LMI0 direct calls), :208-225 (LMIOld in
    MyLmiOracle under Ell and EllStable: < 300 / < 400)"""
    f1 = LMI_TAB[3:15]
    assert _lmi_assess(ref, 2, f1, None, np.zeros(3), 2)[0] == 1                     # test_lmi0_oracle: cut.is_some()
    f0 = [1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0]               # lmi0_matrices()
    assert _lmi_assess(ref, 2, f0, None, [1.0, 0.0, 1.0], 2)[0] == 0                 # feasible
    cut, g, ep = _lmi_assess(ref, 2, f0, None, [-1.0, 0.0, -1.0], 2)                 # infeasible: g = (-1, -0, -0), ep = 1
    assert cut == 1 and g[0] == -1.0 and g[1] == 0.0 and g[2] == 0.0 and ep == 1.0
    assert np.signbit(g[1]) and np.signbit(g[2])                                     # the reference negates sym_quad: -0.0
    assert _lmi_assess(ref, 2, f0, None, [1.0, 1.0, 1.0], 2)[0] == 1                 # infeasible2
    # LMIOld: same element arithmetic as LMIOracle => the same trajectory, not merely < 300 / < 400
    for kind, bound, exact, gam in ((A.KIND_ELL, 300, 11, -3.7502205782089257), (A.KIND_ELLSTABLE, 400, 21, -3.5724046875091333)):
        o = one(ref, oracle_id=A.ORACLE_LMI_LINOPT, n=3, kind=kind, kappa0=10.0, gamma0=math.inf, table=LMI_TAB, aux=[2, 2, 3, 0, 0, 0, 0, 1])
        assert o.has_xbest[0] == 1 and o.num_iters[0] < bound and o.num_iters[0] == exact and o.gamma[0] == gam
    # as plain feasibility oracles on a random instance: LMIOld == LMI bit for bit; LMI0 runs
    p = W.lmi(6, n=12, m=4)
    a = ref.solve(p)
    p.oracle_id = A.ORACLE_LMI_OLD
    b = ref.solve(p)
    assert np.array_equal(a.num_iters, b.num_iters) and np.array_equal(a.xc, b.xc) and np.array_equal(a.exit_status, b.exit_status)
    o = ref.solve(W.lmi0(6, n=12, m=4))
    assert set(o.exit_status.tolist()) <= {A.EXIT_FEASIBLE, A.EXIT_NOSOLN, A.EXIT_TOL, A.EXIT_MAXIT}
    assert (o.exit_status == A.EXIT_FEASIBLE).any()


def test_lowpass(ref):  # src/oracles/lowpass_oracle.rs:153-194 literal case is degenerate
    t = ref.lowpass_table(32)
    p = W.lowpass(1, t)
    p.params = ref.lowpass_case_params()
    p.gamma0 = np.array([p.params[4]])
    o = ref.solve(p)
    assert (o.num_iters[0], o.exit_status[0], o.has_xbest[0]) == (0, A.EXIT_NOSOLN, 0)
    o = ref.solve(W.lowpass(1, t))  # survey-derived: sane bounds → 12481
    assert (o.num_iters[0], o.exit_status[0]) == (12481, A.EXIT_TOL)
    assert o.gamma[0] == 0.0004135142518411313


# ---------------------------------------------------------------- §4(a): EllCalc known answers
def test_ell_calc_core(ref):  # src/ell_calc.rs:943-968
    _, (cst1, cst2, half_n) = ref.calc(4, 5)
    assert half_n == 2.0 and cst1 == pytest.approx(16 / 15, **APPROX) and cst2 == pytest.approx(0.4, **APPROX)
    _, v = ref.calc(4, 3, 1.0, 2.0, 4.0, 2.0, 12.0)
    assert v == pytest.approx((1.2, 0.8, 0.8), **APPROX)
    _, v = ref.calc(4, 4, 1.0, 2.0, 6.0)
    assert v == pytest.approx((1.2, 0.8, 0.8), **APPROX)


def test_ell_calc(ref):  # src/ell_calc.rs:971-1046, 1120-1186
    st, v = ref.calc(4, 1, tsq=0.01)
    assert st == A.SUCCESS and v == pytest.approx((0.02, 0.4, 16 / 15), **APPROX)
    st, v = ref.calc(4, 0, 0.05, tsq=0.01)
    assert st == A.SUCCESS and v == pytest.approx((0.06, 0.8, 0.8), **APPROX)
    assert ref.calc(4, 0, 0.11, tsq=0.01)[0] == A.NOSOLN
    assert ref.calc(4, 0, 0.0, tsq=0.01)[1] == pytest.approx((0.02, 0.4, 16 / 15), **APPROX)
    assert ref.calc(4, 2, -0.05, tsq=0.01)[0] == A.NOEFFECT
    st, v = ref.calc(4, 1, 0.0, 0.11, tsq=0.01)  # parallel central falls back
    assert st == A.SUCCESS and v == pytest.approx((0.02, 0.4, 16 / 15), **APPROX)
    st, v = ref.calc(4, 1, 0.0, 0.05, tsq=0.01)
    assert st == A.SUCCESS and v == pytest.approx((0.02, 0.8, 1.2), **APPROX)
    assert ref.calc(4, 0, 0.07, 0.03, tsq=0.01)[0] == A.NOSOLN
    st, v = ref.calc(4, 0, 0.0, 0.05, tsq=0.01)
    assert st == A.SUCCESS and v == pytest.approx((0.02, 0.8, 1.2), **APPROX)
    st, v = ref.calc(4, 0, 0.05, 0.11, tsq=0.01)
    assert st == A.SUCCESS and v == pytest.approx((0.06, 0.8, 0.8), **APPROX)
    st, v = ref.calc(4, 0, 0.01, 0.04, tsq=0.01)
    assert st == A.SUCCESS and v == pytest.approx((0.0232, 0.928, 1.232), **APPROX)
    assert ref.calc(4, 2, -0.07, 0.07, tsq=0.01)[0] == A.NOEFFECT
    assert ref.calc(4, 2, -0.04, 0.0625, tsq=0.01)[0] == A.NOEFFECT
    assert ref.calc(4, 2, 0.07, 0.03, tsq=0.01)[0] == A.NOSOLN  # calc_parallel_q beta1 < beta0


def test_arr_ops_on_the_path(ref):  # src/arr.rs:676-709: dot_mv, dot, linspace known answers
    L = ref.lib()
    m, v, y = np.eye(3), np.array([2.0, 3.0, 4.0]), np.zeros(3)
    L.ellref_dot_mv(m.ctypes.data, v.ctypes.data, 3, 3, y.ctypes.data)
    assert y[0] == 2.0 and np.array_equal(y, v)
    a, b = np.array([1.0, 2.0, 3.0]), np.array([4.0, 5.0, 6.0])
    assert L.ellref_dot(a.ctypes.data, b.ctypes.data, 3) == 32.0
    out = np.zeros(5)
    L.ellref_linspace(0.0, 10.0, 5, out.ctypes.data)
    assert out[2] == 5.0 and out[0] == 0.0 and out[4] == 10.0
    # sequential, separately rounded: 1e16 + 1 + 1 stays 1e16 left to right (a pairwise or fused sum would not)
    a, b = np.array([1e16, 1.0, 1.0, -1e16]), np.ones(4)
    assert L.ellref_dot(a.ctypes.data, b.ctypes.data, 4) == 0.0


def test_ell_calc_additional_tests(ref):  # src/ell_calc_additional_tests.rs:6-88 — exact assert_eq! on (rho, sigma, delta)
    assert ref.calc(4, 1, tsq=0.0) == (A.SUCCESS, (0.0, 0.4, 16.0 / 15.0))                      # central cut, tsq = 0
    st, v = ref.calc(4, 1, tsq=0.01)
    assert st == A.SUCCESS and min(v) > 0.0
    assert ref.calc(4, 1, 0.0, 0.0, tsq=0.0) == (A.SUCCESS, (0.0, 0.4, 16.0 / 15.0))            # parallel central (0, 0)
    st, v = ref.calc(4, 1, 0.0, 0.05, tsq=0.01)
    assert st == A.SUCCESS and min(v) > 0.0
    assert ref.calc(4, 1, 0.0, -0.1, tsq=0.01) == (A.NOSOLN, (0.0, 0.0, 0.0))
    st, v = ref.calc(4, 1, 0.0, 0.1, tsq=0.2)
    assert st == A.SUCCESS and min(v) > 0.0
    assert ref.calc(4, 1, 0.0, -0.1, tsq=0.2) == (A.NOSOLN, (0.0, 0.0, 0.0))


def test_ell_construct_and_accessors(ref):  # src/ell_test.rs:12-20, 131-155
    s = ref.Space(A.KIND_ELL, 4, kappa=0.01)
    st = s.state()
    assert st["kappa"] == 0.01 and st["tsq"] == 0.0 and np.array_equal(st["mq"], np.eye(4)) and np.array_equal(st["xc"], np.zeros(4))
    s = ref.Space(A.KIND_ELL, 3, kappa=1.0, diag=np.array([1.0, 2.0, 3.0]))
    st = s.state()
    assert st["kappa"] == 1.0 and np.array_equal(st["mq"], np.diag([1.0, 2.0, 3.0])) and st["tsq"] == 0.0
    s = ref.Space(A.KIND_ELL, 2, kappa=1.0)
    s.set_xc(np.array([1.0, 2.0]))
    assert np.array_equal(s.state()["xc"], [1.0, 2.0]) and s.state()["tsq"] == 0.0


# ---------------------------------------------------------------- §4(a): Ell / EllStable single steps
@pytest.mark.parametrize("kind", [A.KIND_ELL, A.KIND_ELLSTABLE])
def test_space_single_steps(ref, kind):  # src/ell.rs:246-314, src/ell_stable.rs:225-298
    g = 0.5 * np.ones(4)
    mk = lambda: ref.Space(kind, 4, kappa=0.01)  # noqa: E731
    s = mk()
    assert s.update(1, g, 0.0) == A.SUCCESS
    st = s.state()
    assert np.array_equal(st["xc"], -0.01 * np.ones(4))
    if kind == A.KIND_ELL:
        assert np.array_equal(st["mq"], np.eye(4) - 0.1 * np.ones((4, 4)))
    assert st["kappa"] == pytest.approx(0.16 / 15, **APPROX) and st["tsq"] == pytest.approx(0.01, **APPROX)
    s = mk()
    assert s.update(0, g, 0.05) == A.SUCCESS
    st = s.state()
    assert st["xc"][0] == pytest.approx(-0.03, **APPROX) and st["kappa"] == pytest.approx(0.008, **APPROX)
    if kind == A.KIND_ELL:
        assert st["mq"][0, 0] == pytest.approx(0.8, **APPROX)
    s = mk()
    assert s.update(1, g, 0.0, 0.05) == A.SUCCESS
    st = s.state()
    assert np.array_equal(st["xc"], -0.01 * np.ones(4)) and st["kappa"] == pytest.approx(0.012, **APPROX)
    if kind == A.KIND_ELL:
        assert np.array_equal(st["mq"], np.eye(4) - 0.2 * np.ones((4, 4)))
    s = mk()
    assert s.update(0, g, 0.01, 0.04) == A.SUCCESS
    st = s.state()
    assert st["xc"][0] == pytest.approx(-0.0116, **APPROX) and st["kappa"] == pytest.approx(0.01232, **APPROX)
    if kind == A.KIND_ELL:
        assert st["mq"][0, 0] == pytest.approx(1 - 0.232, **APPROX)
    s = mk()  # ParallelCut(-0.04, Some(0.0625)): bias → Success with no change, q → NoEffect
    assert s.update(0, g, -0.04, 0.0625) == A.SUCCESS
    st = s.state()
    assert np.allclose(st["xc"], 0.0, atol=1e-6) and st["kappa"] == pytest.approx(0.01, **APPROX)
    if kind == A.KIND_ELL:
        assert np.allclose(st["mq"], np.eye(4), atol=1e-6)
    s = mk()
    assert s.update(2, g, -0.04, 0.0625) == A.NOEFFECT
    st = s.state()
    assert np.array_equal(st["xc"], np.zeros(4)) and st["kappa"] == 0.01


def test_ell_no_defer_trick(ref):  # src/ell.rs:341-354
    s = ref.Space(A.KIND_ELL, 4, kappa=0.01, no_defer_trick=True)
    assert s.update(1, 0.5 * np.ones(4), 0.0) == A.SUCCESS
    st = s.state()
    assert st["kappa"] == 1.0
    plain = ref.Space(A.KIND_ELL, 4, kappa=0.01)
    plain.update(1, 0.5 * np.ones(4), 0.0)
    p = plain.state()
    assert np.array_equal(st["mq"], p["mq"] * p["kappa"])


# ---------------------------------------------------------------- §4(a): LDLTMgr
def _ldlt(ref, m, semi=False):
    m = np.array(m, dtype=float)
    n = m.shape[0]
    pos, wit, ep = np.zeros(2, np.int64), np.zeros(n), C.c_double()
    spd = ref.lib().ellref_ldlt(n, m.ctypes.data, int(semi), pos.ctypes.data, wit.ctypes.data, C.byref(ep))
    return bool(spd), tuple(pos), wit, ep.value


def test_ldlt(ref):  # src/oracles/ldlt_mgr.rs:182-268
    assert _ldlt(ref, [[25, 15, -5], [15, 18, 0], [-5, 0, 11]])[0]
    spd, pos, wit, ep = _ldlt(ref, [[18, 22, 54, 42], [22, -70, 86, 62], [54, 86, -174, 134], [42, 62, 134, -106]])
    assert not spd and pos == (0, 2)
    spd, pos, wit, ep = _ldlt(ref, [[0, 15, -5], [15, 18, 0], [-5, 0, 11]])
    assert not spd and pos == (0, 1) and wit[0] == 1.0 and ep == 0.0
    # semidefinite variant (ldlt_mgr.rs:230-254) — off the hot path, pinned because it is cheap
    assert _ldlt(ref, [[0, 15, -5], [15, 18, 0], [-5, 0, 11]], semi=True)[0]
    spd, pos, wit, ep = _ldlt(ref, [[0, 15, -5], [15, 18, 0], [-5, 0, -20]], semi=True)
    assert not spd and ep == pytest.approx(20.0, **APPROX)
    assert not _ldlt(ref, [[0, 15, -5], [15, 18, 0], [-5, 0, 20]])[0]
    assert _ldlt(ref, [[0, 15, -5], [15, 18, 0], [-5, 0, 20]], semi=True)[0]


# ---------------------------------------------------------------- §8f rank 4: GEMV-sharing siblings
POWER_A = np.array([[3.7, -3.6, 0.7], [-3.6, 4.3, -2.8], [0.7, -2.8, 5.4]])


def test_power_iteration_pins(ref):  # src/power_iteration.rs:178-229
    x0 = [0.3, 0.5, 0.4]
    assert ref.power_iteration(1, POWER_A, x0, 2000, 1e-7)[1] == 22
    assert ref.power_iteration(4, POWER_A, x0, 2000, 1e-7)[1] == 21
    assert ref.power_iteration(2, POWER_A, x0, 2000, 1e-14)[1] == 23
    assert ref.power_iteration(3, POWER_A, x0, 2000, 1e-14)[1] == 23
    assert ref.power_iteration(4, POWER_A, x0, 5, 1e-14)[1] == 5


def test_conjugate_gradient_pins(ref):  # src/conjugate_gradient.rs:86-124
    a = np.array([[4.0, 1.0], [1.0, 3.0]])
    for x0 in (None, [1.0, 1.0]):
        x, it = ref.conjugate_gradient(a, [1.0, 2.0], x0, 1e-5, 1000)
        assert it >= 0 and x == pytest.approx([0.0909091, 0.6363636], abs=1e-5)
    assert ref.conjugate_gradient(np.zeros((2, 2)), [1.0, 1.0], None, 1e-5, 10)[1] < 0
    x, it = ref.conjugate_gradient(a, [1.0, 2.0], None, 1e-10, 1000)
    r = np.array([1.0, 2.0]) - a @ x
    assert it >= 0 and math.sqrt(r @ r) < 1e-10


def test_host_linalg_mirror_equals_checker(ref):
    """ellalgo_rs_b200/linalg.py (the host loops that drive the device GEMV) against the restatement, with a host matrix"""
    from ellalgo_rs_b200 import linalg as LA

    class HostMat:
        def __init__(self, a): self.a = np.asarray(a, float)
        def dot_mv(self, x): return np.array([float(np.cumsum(r * x)[-1]) for r in self.a])
    x0 = np.array([0.3, 0.5, 0.4])
    for v, f, tol in ((1, LA.power_iteration, 1e-7), (4, LA.power_iteration4, 1e-7), (2, LA.power_iteration2, 1e-14),
                      (3, LA.power_iteration3, 1e-14)):
        e, it, x = f(HostMat(POWER_A), x0, LA.Options(2000, tol))
        er, itr, xr = ref.power_iteration(v, POWER_A, x0, 2000, tol)
        assert (it, e) == (itr, er) and np.array_equal(x, xr)
    with pytest.raises(LA.ConjGradError):
        LA.conjugate_gradient(HostMat(np.zeros((2, 2))), [1.0, 1.0], None, 1e-5, 10)
