"""GPU parity tests, batched mode: the CUDA path (through the C ABI, ellcuda_batch_solve) against the CPU checker
on the same seeded inputs."""
import math

import numpy as np
import pytest

from ellalgo_rs_b200 import _abi as A
from ellalgo_rs_b200 import batch
from ellalgo_rs_b200 import workloads as W
from ellalgo_rs_b200.problem import BatchProblem
from parity import assert_parity

pytestmark = pytest.mark.gpu


def both(ref, prob):
    return batch.solve_batch(prob), ref.solve(prob)


# ------------------------------------------------------------------ device exp/ln == host libm
def test_device_exp_ln_equal_glibc(ref):
    """The oracles' exp/ln on the device (csrc/glibc_math.cuh) must equal the host's libm — what Rust's f64::exp /
    f64::ln call (src/oracles/profit_oracle.rs:54-113) — in every bit, special values included."""
    import ctypes as C
    from ellalgo_rs_b200._lib import check, lib
    rng = np.random.default_rng(11)
    n = 1 << 20
    x = np.concatenate([rng.uniform(-10, 10, n), rng.uniform(-745, 710, n), rng.uniform(-1e-3, 1e-3, n),
                        rng.integers(0, 2**64, n, dtype=np.uint64).view(np.float64),
                        [0.0, -0.0, np.inf, -np.inf, np.nan, 709.782712893384, -745.1332191019412, 1e-320]])
    y = np.empty_like(x)
    check(lib().ellcuda_math_probe(0, x.ctypes.data, y.ctypes.data, x.size, 0))
    want = ref.libm(0, x)      # numpy's SIMD exp is not libm's; the checker calls glibc
    same = (y.view(np.uint64) == want.view(np.uint64)) | (np.isnan(y) & np.isnan(want))
    assert same.all(), f"exp differs at {x[~same][:5]}"
    x = np.concatenate([rng.uniform(0.9, 1.1, n), np.exp(rng.uniform(-5, 15, n)), rng.uniform(0, 1e-310, n),
                        rng.integers(0, 2**64, n, dtype=np.uint64).view(np.float64),
                        [0.0, -0.0, 1.0, np.inf, -1.0, np.nan, 5e-324, 1.7976931348623157e308]])
    check(lib().ellcuda_math_probe(1, x.ctypes.data, y.ctypes.data, x.size, 0))
    want = ref.libm(1, x)
    same = (y.view(np.uint64) == want.view(np.uint64)) | (np.isnan(y) & np.isnan(want))
    assert same.all(), f"ln differs at {x[~same][:5]}"


# ------------------------------------------------------------------ n = 2, thread per problem
def test_example1_pinned(ref):  # src/example1.rs:39-50
    g, r = both(ref, W.example1())
    assert g.num_iters[0] == 25 and g.exit_status[0] == A.EXIT_TOL and g.has_xbest[0] == 1
    assert_parity(g, r, exact=True, what="example1")


FIXTURES_N2 = [
    ("example1_rr", dict(oracle_id=A.ORACLE_EXAMPLE1_RR, kappa0=10.0, gamma0=-math.inf, tolerance=1e-10), 25),
    ("example4", dict(oracle_id=A.ORACLE_EXAMPLE4, kappa0=10.0, gamma0=-math.inf, tolerance=1e-10), 82),
    ("example1_infeasible1", dict(oracle_id=A.ORACLE_EXAMPLE1, kappa0=1.0, diag0=[10.0, 10.0], xc0=[100.0, 100.0], gamma0=-math.inf), None),
    ("example1_infeasible2", dict(oracle_id=A.ORACLE_EXAMPLE1, kappa0=1.0, diag0=[10.0, 10.0], gamma0=100.0), None),
    ("halfplane", dict(oracle_id=A.ORACLE_HALFPLANE, mode=A.MODE_FEAS, kappa0=10.0, max_iters=200), 0),
    ("always_cut_feas", dict(oracle_id=A.ORACLE_ALWAYS_CUT, mode=A.MODE_FEAS, kappa0=10.0, max_iters=200), 2),
    ("always_cut_feas5", dict(oracle_id=A.ORACLE_ALWAYS_CUT, mode=A.MODE_FEAS, kappa0=10.0, max_iters=5), 2),
    ("always_cut_optim", dict(oracle_id=A.ORACLE_ALWAYS_CUT, kappa0=10.0, gamma0=0.0, max_iters=5), 2),
    ("always_cut_q", dict(oracle_id=A.ORACLE_ALWAYS_CUT, mode=A.MODE_OPTIM_Q, kappa0=10.0, gamma0=0.0, max_iters=5), 2),
    ("optim_box", dict(oracle_id=A.ORACLE_OPTIM_BOX, kappa0=10.0, gamma0=0.0, max_iters=200), None),
    ("optim_box_nosoln", dict(oracle_id=A.ORACLE_OPTIM_BOX, kappa0=10.0, gamma0=100.0, max_iters=4), 0),
    ("optim_box_q", dict(oracle_id=A.ORACLE_OPTIM_BOX, mode=A.MODE_OPTIM_Q, kappa0=10.0, gamma0=0.0, max_iters=200), None),
    ("optim_box_q_nosoln", dict(oracle_id=A.ORACLE_OPTIM_BOX, mode=A.MODE_OPTIM_Q, kappa0=10.0, gamma0=100.0, max_iters=20), 0),
    ("example2", dict(oracle_id=A.ORACLE_EXAMPLE2, mode=A.MODE_FEAS, kappa0=10.0), 1),
    ("example2_inf", dict(oracle_id=A.ORACLE_EXAMPLE2, mode=A.MODE_FEAS, kappa0=10.0, xc0=[100.0, 100.0]), 0),
    ("example3", dict(oracle_id=A.ORACLE_EXAMPLE3, mode=A.MODE_FEAS, kappa0=100.0, params=[-20.0], tolerance=1e-8), None),
    ("max_iters_0", dict(oracle_id=A.ORACLE_EXAMPLE1, kappa0=10.0, gamma0=-math.inf, max_iters=0), 0),
]


@pytest.mark.parametrize("name,kw,pinned", FIXTURES_N2, ids=[f[0] for f in FIXTURES_N2])
@pytest.mark.parametrize("kind", [A.KIND_ELL, A.KIND_ELLSTABLE])
def test_fixtures_n2(ref, name, kw, pinned, kind):
    prob = BatchProblem(n=2, count=3, kind=kind, **kw)  # 3 identical copies: also exercises the partial warp
    g, r = both(ref, prob)
    if pinned is not None and kind == A.KIND_ELL:
        assert (g.num_iters == pinned).all(), f"{name}: {g.num_iters} != pinned {pinned}"
    assert_parity(g, r, exact=True, what=name)


@pytest.mark.parametrize("kind", [A.KIND_ELL, A.KIND_ELLSTABLE])
def test_quasicvx(ref, kind):  # src/quasicvx.rs:61-131 (device exp == glibc exp, so bit-identical)
    prob = BatchProblem(oracle_id=A.ORACLE_QUASICVX, n=2, count=1, kind=kind, kappa0=1.0, diag0=[10.0, 10.0], gamma0=0.0,
                        tolerance=1e-8)
    g, r = both(ref, prob)
    if kind == A.KIND_ELL:
        assert g.num_iters[0] == 35
    assert_parity(g, r, exact=True, what="quasicvx")


def test_profit_pinned_and_random(ref):  # src/oracles/profit_oracle.rs:207-223 + C2 generator
    prob = W.profit(200000)
    g, r = both(ref, prob)
    assert g.num_iters[0] == 83 and g.xbest[0, 0] <= math.log(30.5)
    assert_parity(g, r, exact=True, what="profit")


def test_profit_rb_and_q(ref):  # :225-260
    p = W.profit(1)
    p.oracle_id = A.ORACLE_PROFIT_RB
    p.params = np.concatenate([W.PROFIT_REF, [0.003, 0.007, 1.0, 1.0, 1.0]])
    g, r = both(ref, p)
    assert g.num_iters[0] == 90
    assert_parity(g, r, exact=True, what="profit_rb")
    p = W.profit(4096)
    p.oracle_id, p.mode = A.ORACLE_PROFIT_Q, A.MODE_OPTIM_Q
    g, r = both(ref, p)
    assert g.num_iters[0] == 29
    assert_parity(g, r, exact=True, what="profit_q")


def test_bsearch_on_device(ref):  # src/example3.rs:68-85: bsearch + BSearchAdaptor, whole nested loop on the device
    prob = BatchProblem(oracle_id=A.ORACLE_EXAMPLE3, n=2, count=1, mode=A.MODE_BSEARCH, kappa0=100.0, gamma0=-100.0,
                        params=[100.0, 1e-8, 2000.0], tolerance=1e-8)
    g, r = both(ref, prob)
    assert (g.has_xbest[0], g.num_iters[0]) == (1, 34)
    assert_parity(g, r, exact=True, what="bsearch example3")
    rng = np.random.default_rng(5)
    cnt = 500
    params = np.stack([rng.uniform(-50, 150, cnt), 10.0 ** rng.uniform(-9, -3, cnt), rng.integers(5, 60, cnt).astype(float)], axis=1)
    for kind in (A.KIND_ELL, A.KIND_ELLSTABLE):
        prob = BatchProblem(oracle_id=A.ORACLE_EXAMPLE3, n=2, count=cnt, mode=A.MODE_BSEARCH, kind=kind,
                            kappa0=rng.uniform(1.0, 200.0, cnt), xc0=rng.uniform(-3, 3, (cnt, 2)),
                            gamma0=rng.uniform(-150, -60, cnt), params=params, tolerance=1e-8, max_iters=300)
        g, r = both(ref, prob)
        assert_parity(g, r, exact=True, what=f"bsearch random kind={kind}")


def test_flags_n2(ref):
    for kw in (dict(no_defer_trick=True), dict(use_parallel_cut=False)):
        prob = BatchProblem(oracle_id=A.ORACLE_EXAMPLE4, n=2, count=2, kappa0=10.0, gamma0=-math.inf, tolerance=1e-10, **kw)
        g, r = both(ref, prob)
        assert_parity(g, r, exact=True, what=str(kw))


# ------------------------------------------------------------------ 3 <= n <= 32, warp per problem
@pytest.mark.parametrize("n,count,spread", [(32, 300, 1.0), (32, 64, 3.0), (5, 33, 2.0), (17, 40, 1.0), (3, 7, 1.0)])
def test_bowl(ref, n, count, spread):  # tests/integration_test.rs:90-105 — no libm: bit-identical
    prob = W.bowl(count, n=n, spread=spread, max_iters=20000)
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what=f"bowl n={n}")


def test_bowl_reference_case(ref):  # tests/integration_test.rs:107-131: n=5, target 1..5, kappa 10
    prob = BatchProblem(oracle_id=A.ORACLE_BOWL, n=5, count=1, kappa0=10.0, gamma0=math.inf, params=[1.0, 2, 3, 4, 5],
                        max_iters=2000, tolerance=1e-10)
    g, r = both(ref, prob)
    assert g.has_xbest[0] == 1 and math.sqrt(((g.xbest[0] - np.arange(1, 6)) ** 2).mean()) < 3.0
    assert_parity(g, r, exact=True, what="HighDimOracle")


def test_bowl_flags(ref):
    for kw in (dict(no_defer_trick=True), dict(max_iters=2000), dict(max_iters=1)):
        prob = W.bowl(16, n=32, max_iters=kw.get("max_iters", 500))
        prob.no_defer_trick = kw.get("no_defer_trick", False)
        g, r = both(ref, prob)
        assert_parity(g, r, exact=True, what=str(kw))


def test_bowl_matrix_start(ref):  # Ell::new (diag) and new_with_matrix (symmetric dense) starts
    n, count = 32, 8
    rng = np.random.default_rng(7)
    prob = W.bowl(count, n=n, max_iters=300)
    prob.kappa0 = np.array([1.0])
    prob.diag0 = rng.uniform(50.0, 150.0, size=(count, n))
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what="diag start")
    a = rng.uniform(-1, 1, size=(n, n))
    prob = W.bowl(count, n=n, max_iters=300)
    prob.kappa0 = np.array([1.0])
    prob.mq0 = 100.0 * np.eye(n) + (a + a.T)
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what="matrix start")


def test_lowpass(ref):  # src/oracles/lowpass_oracle.rs — table built with glibc cos on the host: bit-identical
    t = batch.lowpass_table(32)
    assert np.array_equal(t, ref.lowpass_table(32))
    prob = W.lowpass(12, t)
    g, r = both(ref, prob)
    assert g.num_iters[0] == 12481 and g.exit_status[0] == A.EXIT_TOL  # survey-derived canonical instance
    assert_parity(g, r, exact=True, what="lowpass n=32")


def test_lowpass_literal_case(ref):  # create_lowpass_case(32) is degenerate: NoSoln at iteration 0
    t = batch.lowpass_table(32)
    prob = W.lowpass(1, t)
    prob.params = ref.lowpass_case_params()
    prob.gamma0 = np.array([prob.params[4]])
    g, r = both(ref, prob)
    assert (g.num_iters[0], g.exit_status[0], g.has_xbest[0]) == (0, A.EXIT_NOSOLN, 0)
    assert_parity(g, r, exact=True, what="lowpass literal")


@pytest.mark.parametrize("n", [8, 15])
def test_lowpass_small_n(ref, n):
    t = batch.lowpass_table(n)
    prob = W.lowpass(5, t, n=n, max_iters=5000)
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what=f"lowpass n={n}")


def test_empty_batch_and_bad_args(ref):
    out = batch.solve_batch(W.bowl(0, n=32))
    assert out.num_iters.size == 0
    from ellalgo_rs_b200._lib import EllCudaError
    with pytest.raises(EllCudaError):
        batch.solve_batch(BatchProblem(oracle_id=A.ORACLE_PROFIT, n=3, count=1, params=W.PROFIT_REF))
    odd = W.bowl(2, n=4, max_iters=40)          # a non-symmetric new_with_matrix start is legal (src/ell.rs:31-41)
    odd.mq0 = np.eye(4) + np.arange(16.0).reshape(4, 4) / 64.0
    g, r = both(ref, odd)
    assert_parity(g, r, exact=True, what="non-symmetric mq0")


# ------------------------------------------------------------------ EllStable + LMI (n <= 64), warp per problem
def test_lmi_reference_matrices(ref):  # tests/lmi_tests.rs:14-64: (F1,B1) is feasible at 0; also a far-away start
    f1 = np.array([-7, -11, -11, 3, 7, -18, -18, 8, -2, -8, -8, 1], dtype=float)
    b1 = np.array([33, -9, -9, 26], dtype=float)
    tab = np.concatenate([f1, b1])
    for xc0 in ([0.0, 0.0, 0.0], [30.0, -20.0, 10.0], [1.0, 1.0, 1.0]):
        prob = BatchProblem(oracle_id=A.ORACLE_LMI, n=3, count=2, mode=A.MODE_FEAS, kind=A.KIND_ELLSTABLE, kappa0=10.0,
                            xc0=xc0, table=tab, aux=[2])
        g, r = both(ref, prob)
        assert_parity(g, r, exact=True, what=f"lmi 2x2 from {xc0}")
    g, _ = both(ref, BatchProblem(oracle_id=A.ORACLE_LMI, n=3, count=1, mode=A.MODE_FEAS, kind=A.KIND_ELLSTABLE,
                                  kappa0=10.0, table=tab, aux=[2]))
    assert (g.has_xbest[0], g.num_iters[0], g.exit_status[0]) == (1, 0, A.EXIT_FEASIBLE)


@pytest.mark.parametrize("n,m,count,general", [(64, 16, 64, False), (20, 7, 33, False), (12, 4, 40, True)])
def test_lmi_old_and_lmi0_device_oracles(ref, n, m, count, general):
    """ELL_ORACLE_LMI_OLD (src/oracles/lmi_old_oracle.rs) and ELL_ORACLE_LMI0 (lmi0_oracle.rs) on the device against the
    checker's literal restatements (LMIOld there really forms the full matrix, k outermost)"""
    prob = W.lmi(count, n=n, m=m, general=general, upper_scale=1e-3)
    prob.oracle_id = A.ORACLE_LMI_OLD
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what=f"lmi_old n={n} m={m}")
    prob.oracle_id = A.ORACLE_LMI
    g2 = batch.solve_batch(prob)
    assert np.array_equal(g.num_iters, g2.num_iters) and np.array_equal(g.xc, g2.xc)      # same trajectory as LMIOracle
    prob = W.lmi0(count, n=n, m=m)
    if general:
        prob.mq0 = W.lmi(count, n=n, m=m, general=True, upper_scale=1e-3).mq0
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what=f"lmi0 n={n} m={m}")
    assert len(set(g.exit_status.tolist())) >= 1 and (g.num_iters > 0).any()


def test_lmi0_reference_matrices(ref):  # tests/lmi_tests.rs:79-113: lmi0_matrices(), feasible at (1,0,1)
    f0 = np.array([1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0])
    for xc0, feasible_at_start in (([1.0, 0.0, 1.0], True), ([-1.0, 0.0, -1.0], False), ([1.0, 1.0, 1.0], False)):
        prob = BatchProblem(oracle_id=A.ORACLE_LMI0, n=3, count=2, mode=A.MODE_FEAS, kind=A.KIND_ELLSTABLE, kappa0=10.0,
                            xc0=xc0, table=f0, aux=[2])
        g, r = both(ref, prob)
        assert_parity(g, r, exact=True, what=f"lmi0 2x2 from {xc0}")
        assert (g.num_iters[0] == 0 and g.exit_status[0] == A.EXIT_FEASIBLE) == feasible_at_start


@pytest.mark.parametrize("n,m,count", [(64, 16, 96), (20, 7, 40), (33, 16, 17), (64, 5, 33)])
def test_lmi_diagonal_start(ref, n, m, count):  # C4a: EllStable::new_with_scalar — the matrix stays diagonal
    prob = W.lmi(count, n=n, m=m)
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what=f"lmi diag n={n} m={m}")
    assert (g.exit_status == A.EXIT_FEASIBLE).any()


def test_lmi_diag_vector_start(ref):  # EllStable::new(val, xc)
    prob = W.lmi(24, n=40, m=9)
    prob.kappa0 = np.array([1.0])
    prob.diag0 = np.random.default_rng(3).uniform(50, 150, size=(24, 40))
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what="lmi diag vector start")


@pytest.mark.parametrize("n,m,count,scale", [(64, 16, 24, 0.1), (64, 16, 40, 1e-3), (64, 16, 40, 1e-6), (17, 6, 20, 0.05),
                                             (40, 16, 9, 1e-4)])
def test_lmi_general_start(ref, n, m, count, scale):  # C4b: new_with_matrix with a non-zero strict upper triangle
    prob = W.lmi(count, n=n, m=m, general=True, upper_scale=scale)
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what=f"lmi general n={n} m={m}")


def test_lmi_bad_arguments():
    from ellalgo_rs_b200._lib import EllCudaError
    prob = W.lmi(2, n=8, m=4)
    prob.table = prob.table.copy()
    prob.table[1] += 1.0            # F_0 no longer symmetric
    with pytest.raises(EllCudaError):
        batch.solve_batch(prob)
    with pytest.raises(EllCudaError):
        batch.solve_batch(W.lmi(2, n=65, m=4))


# ------------------------------------------------------------------ edge cases
def test_edge_cases_warp_and_lmi_kernels(ref):
    t = batch.lowpass_table(32)
    for prob in (W.bowl(1, n=32, max_iters=50), W.bowl(9, n=32, max_iters=0), W.lowpass(1, t, max_iters=3),
                 W.lowpass(3, t, max_iters=0), W.lmi(1, n=64, m=16), W.lmi(5, n=64, m=16, general=True, upper_scale=1e-3)):
        g, r = both(ref, prob)
        assert_parity(g, r, exact=True, what=f"edge oracle={prob.oracle_id} count={prob.count} max_iters={prob.max_iters}")
    prob = W.lmi(3, n=12, m=4)
    prob.max_iters = 0
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what="lmi max_iters=0")


def test_lowpass_flags_and_per_instance_kappa(ref):
    t = batch.lowpass_table(32)
    for kw in (dict(no_defer_trick=True), dict(use_parallel_cut=False)):
        prob = W.lowpass(4, t, max_iters=3000)
        for k, v in kw.items():
            setattr(prob, k, v)
        prob.kappa0 = np.array([40.0, 10.0, 80.0, 25.0])
        g, r = both(ref, prob)
        assert_parity(g, r, exact=True, what=f"lowpass {kw}")


def _lowpass_band_cases(t):
    """band layouts that stress the 32-row rounds of the TMEM kernel: bands shorter than a round, an empty transition
    band (wstop == wpass), pass-band of one row, band edges on and next to multiples of 32, a stop-band of 2 rows"""
    prob = W.lowpass(9, t, max_iters=1500)
    p = prob.params
    edges = [(0.001, 0.002), (0.12, 0.12), (31.5 / 479, 63.5 / 479), (32.5 / 479, 64.5 / 479), (0.30, 0.31),
             (0.05, 0.995), (0.90, 0.9979), (63.5 / 479, 95.5 / 479), (0.49, 0.51)]
    for i, (wp, ws) in enumerate(edges):
        p[i, 0], p[i, 1] = wp, ws
    return prob


@pytest.mark.parametrize("variant", ["0", "1", "2"])
def test_lowpass32_tmem_kernel_band_layouts_and_initial_matrices(ref, variant, monkeypatch):
    """batch_lowpass.cu: both variants (spectrum in shared memory / tail columns in TMEM) against the checker on awkward
    band layouts, and with Ell::new (diag0) and Ell::new_with_matrix (mq0) starts, which go through the TMEM store path"""
    monkeypatch.setenv("ELLCUDA_LOWPASS_VARIANT", variant)
    t = batch.lowpass_table(32)
    g, r = both(ref, _lowpass_band_cases(t))
    assert_parity(g, r, exact=True, what=f"lowpass band layouts, variant {variant}")
    rng = np.random.default_rng(7)
    prob = W.lowpass(5, t, max_iters=800)
    prob.kappa0 = np.full(5, 1.0)
    prob.diag0 = 20.0 + 40.0 * rng.random((5, 32))
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what=f"lowpass diag0, variant {variant}")
    prob = W.lowpass(3, t, max_iters=800)
    a = rng.standard_normal((3, 32, 32)) * 0.1
    spd = np.einsum("pij,pkj->pik", a, a) + np.eye(32)[None]
    prob.mq0 = np.ascontiguousarray(0.5 * (spd + spd.transpose(0, 2, 1)))     # exactly symmetric (the C ABI checks it)
    prob.kappa0 = np.full(3, 30.0)
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what=f"lowpass mq0, variant {variant}")


def test_lowpass32_more_problems_than_resident_warps(ref):
    """the persistent kernel's queue: 148 SMs x 16 warps = 2368 resident problems; 2600 short problems must all come back,
    equal to the checker's first 64 and to a re-run"""
    t = batch.lowpass_table(32)
    prob = W.lowpass(2600, t, max_iters=40)
    g1 = batch.solve_batch(prob)
    g2 = batch.solve_batch(prob)
    assert np.array_equal(g1.num_iters, g2.num_iters) and np.array_equal(g1.xc, g2.xc)
    assert (g1.exit_status > 0).all()
    small = W.lowpass(64, t, max_iters=40)
    r = ref.solve(small)
    assert np.array_equal(g1.num_iters[:64], r.num_iters) and np.array_equal(g1.xc[:64], r.xc)


def test_non_finite_inputs_propagate_like_the_reference(ref):
    """no guard is added that the reference does not have: NaN / inf centres give the same status"""
    prob = W.bowl(4, n=32, max_iters=50)
    prob.xc0 = np.zeros((4, 32))
    prob.xc0[1, 3] = np.nan
    prob.xc0[2, 0] = np.inf
    prob.params[3, :] = 0.0          # target == centre: g = 0, omega = 0, sigma = 0/0 — Success with NaN state, then tsq<tol
    g, r = both(ref, prob)
    assert np.array_equal(g.num_iters, r.num_iters) and np.array_equal(g.exit_status, r.exit_status)
    assert np.array_equal(g.has_xbest, r.has_xbest)
    assert np.array_equal(np.isnan(g.xc), np.isnan(r.xc)) and np.array_equal(np.isnan(g.tsq), np.isnan(r.tsq))


# ------------------------------------------------------------------ BASELINE.json's full sizes, size-independent properties
def test_full_size_profit_one_million(ref):
    """configs[1]: 2^20 profit problems.  Properties: problem 0 is the reference test (83 iterations); a problem's
    result does not depend on its position or on the batch around it (the first 4096 equal the checker's, a duplicated
    block equals the original); re-running gives identical bits."""
    cnt = 1 << 20
    prob = W.profit(cnt)
    prob.params[cnt // 2: cnt // 2 + 4096] = prob.params[:4096]
    g = batch.solve_batch(prob)
    assert g.num_iters[0] == 83 and (g.exit_status == A.EXIT_TOL).all() and g.has_xbest.all()
    r = ref.solve(W.profit(4096))
    for f in ("num_iters", "xbest", "gamma"):
        assert np.array_equal(getattr(g, f)[:4096], getattr(r, f))
        assert np.array_equal(getattr(g, f)[cnt // 2: cnt // 2 + 4096], getattr(r, f))
    g2 = batch.solve_batch(prob)
    assert np.array_equal(g.xbest, g2.xbest) and np.array_equal(g.num_iters, g2.num_iters)


def test_full_size_lowpass_64k_order_independent(ref):
    """configs[2]: 65536 lowpass n=32 problems.  The dynamic problem queue must not leak state between problems:
    solving the batch in reversed order gives the reversed results, bit for bit; the head of the batch equals the checker."""
    t = batch.lowpass_table(32)
    cnt = 1 << 16
    prob = W.lowpass(cnt, t)
    g = batch.solve_batch(prob)
    rev = W.lowpass(cnt, t)
    rev.params = np.ascontiguousarray(prob.params[::-1])
    rev.gamma0 = np.ascontiguousarray(prob.gamma0[::-1])
    gr = batch.solve_batch(rev)
    for f in ("num_iters", "exit_status", "has_xbest", "xbest", "gamma"):
        assert np.array_equal(getattr(g, f), getattr(gr, f)[::-1]), f
    r = ref.solve(W.lowpass(48, t))
    assert np.array_equal(g.num_iters[:48], r.num_iters) and np.array_equal(g.xbest[:48], r.xbest)
    assert g.num_iters[0] == 12481


def test_full_size_bowl_one_million(ref):
    """SURVEY C3': 2^20 bowl problems at n=32 (max_iters capped at 200 for wall time: every problem hits MAXIT)"""
    cnt = 1 << 20
    prob = W.bowl(cnt, n=32, max_iters=200)
    g = batch.solve_batch(prob)
    assert (g.num_iters <= 200).all() and (g.exit_status[g.num_iters == 200] == A.EXIT_MAXIT).all()
    assert (g.num_iters == 200).mean() > 0.99
    early = np.flatnonzero(g.num_iters < 200)[:32]      # the rare early exits are checked against the checker too
    idx = np.r_[0:64, cnt - 64:cnt, early]
    sub = W.bowl(cnt, n=32, max_iters=200)
    sub = BatchProblem(oracle_id=A.ORACLE_BOWL, n=32, count=idx.size, kappa0=100.0, xc0=np.zeros(32), gamma0=math.inf,
                       params=np.ascontiguousarray(prob.params[idx]), max_iters=200, tolerance=1e-10)
    r = ref.solve(sub)
    assert np.array_equal(g.num_iters[idx], r.num_iters) and np.array_equal(g.exit_status[idx], r.exit_status)
    assert np.array_equal(g.xc[idx], r.xc) and np.array_equal(g.gamma[idx], r.gamma) and np.array_equal(g.kappa[idx], r.kappa)


def test_lmi_tests_rs_fixture_on_device(ref):
    """tests/lmi_tests.rs:121-225 (MyLmiOracle over two LMIOracle blocks, cutting_plane_optim) with Ell and EllStable:
    the reference asserts < 300 / < 400 iterations; the checker gives 11 / 21 and these gammas."""
    from test_oracle_pins import LMI_TAB
    for kind, iters, gamma in ((A.KIND_ELL, 11, -3.7502205782089257), (A.KIND_ELLSTABLE, 21, -3.5724046875091333)):
        prob = BatchProblem(oracle_id=A.ORACLE_LMI_LINOPT, n=3, count=2, kind=kind, kappa0=10.0, gamma0=math.inf, table=LMI_TAB,
                            aux=[2, 2, 3])
        g, r = both(ref, prob)
        assert g.has_xbest.all() and (g.num_iters == iters).all() and (g.gamma == gamma).all()
        assert_parity(g, r, exact=True, what=f"lmi_tests.rs kind={kind}")
    rng = np.random.default_rng(2)
    prob = BatchProblem(oracle_id=A.ORACLE_LMI_LINOPT, n=3, count=300, kappa0=rng.uniform(5, 50, 300),
                        xc0=rng.uniform(-1, 1, (300, 3)), gamma0=math.inf, table=LMI_TAB, aux=[2, 2, 3])
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what="lmi_tests.rs random starts")
