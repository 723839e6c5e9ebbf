"""GPU parity tests of the batched `Ell` kernel for 64 < n <= 224 with M in shared memory (packed lower triangle), one CTA per problem
(csrc/batch_ellcta.cu): the quadratic bowl and LowpassOracle under cutting_plane_optim — the reference's own stress test
runs LowpassOracle with 128 taps (tests/stress_tests.rs:6-14).  Bit for bit against the CPU checker."""
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


@pytest.mark.parametrize("n,count,spread,max_iters", [(65, 30, 1.0, 400), (96, 41, 1.0, 500), (100, 23, 3.0, 300), (128, 50, 1.0, 600),
                                                      (160, 12, 1.0, 400), (97, 700, 1.0, 40), (224, 7, 1.0, 300), (193, 5, 2.0, 250)])
def test_bowl_n_above_64(ref, n, count, spread, max_iters):  # tests/integration_test.rs:90-105
    prob = W.bowl(count, n=n, spread=spread, max_iters=max_iters)
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what=f"bowl n={n}")
    assert g.num_iters.max() > 20


def test_bowl_cta_runs_to_tolerance(ref):
    prob = W.bowl(2, n=72, max_iters=60000)
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what="bowl n=72 to tolerance")
    assert (g.exit_status == A.EXIT_TOL).all() and g.num_iters.min() > 10000


@pytest.mark.parametrize("n", [80, 128, 200])
def test_bowl_cta_flags_and_starts(ref, n):
    rng = np.random.default_rng(n)
    for kw in (dict(no_defer_trick=True), dict(use_parallel_cut=False), dict(max_iters=1), dict(max_iters=0)):
        prob = W.bowl(11, n=n, max_iters=kw.get("max_iters", 150))
        prob.no_defer_trick = kw.get("no_defer_trick", False)
        prob.use_parallel_cut = kw.get("use_parallel_cut", True)
        g, r = both(ref, prob)
        assert_parity(g, r, exact=True, what=f"bowl n={n} {kw}")
    prob = W.bowl(9, n=n, max_iters=200)                       # Ell::new(val, xc): per-instance diagonals
    prob.kappa0 = np.ones(1)
    prob.diag0 = rng.uniform(50.0, 150.0, (9, n))
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what=f"bowl n={n} diag start")
    a = rng.uniform(-1, 1, (9, n, n)) / n                      # Ell::new_with_matrix, symmetric dense
    prob.diag0 = None
    prob.mq0 = np.ascontiguousarray(100.0 * (np.eye(n) + a + np.transpose(a, (0, 2, 1))))
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what=f"bowl n={n} matrix start")
    for nd in (True, False):                                   # ... and non-symmetric (src/ell.rs:31-41, 117-128)
        prob.mq0 = np.ascontiguousarray(100.0 * (np.eye(n) + a))
        prob.no_defer_trick = nd
        g, r = both(ref, prob)
        assert_parity(g, r, exact=True, what=f"bowl n={n} non-symmetric matrix start, no_defer={nd}")


@pytest.mark.parametrize("n,count,max_iters", [(128, 4, 1200), (65, 5, 1500), (96, 3, 1000), (160, 2, 500), (224, 2, 300)])
def test_lowpass_n_above_64(ref, n, count, max_iters):  # src/oracles/lowpass_oracle.rs with 128 taps (stress_tests.rs:6-14)
    t = batch.lowpass_table(n)
    prob = W.lowpass(count, t, n=n, max_iters=max_iters)
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what=f"lowpass n={n}")
    assert g.num_iters.max() == max_iters or (g.exit_status != A.EXIT_MAXIT).any()


def test_lowpass_cta_band_layouts(ref):
    """short bands, an empty transition band, band edges on multiples of the CTA's round length"""
    n = 128
    t = batch.lowpass_table(n)
    prob = W.lowpass(5, t, n=n, max_iters=400)
    m1 = 15 * n - 1
    for i, (wp, ws) in enumerate([(0.001, 0.002), (0.12, 0.12), (127.5 / m1, 255.5 / m1), (0.05, 0.995), (0.90, 0.9979)]):
        prob.params[i, 0], prob.params[i, 1] = wp, ws
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what="lowpass n=128 band layouts")


def test_stress_lowpass_high_dimension_batched(ref):
    """tests/stress_tests.rs:6-14 as a batch: create_lowpass_case(128), Ell::new_with_scalar(1.0, 0), Options(50000, 1e-14)"""
    n = 128
    t = batch.lowpass_table(n)
    p = ref.lowpass_case_params()
    prob = BatchProblem(oracle_id=A.ORACLE_LOWPASS, n=n, count=3, kappa0=1.0, xc0=np.zeros(n), gamma0=p[4], params=p, table=t,
                        aux=[15 * n], max_iters=50000, tolerance=1e-14)
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what="lowpass literal n=128")


def test_cta_more_problems_than_resident_ctas_and_limits(ref):
    prob = W.bowl(148 * 2 + 7, n=120, max_iters=12)
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what="bowl n=120, more than one wave")
    assert batch.solve_batch(W.bowl(0, n=128)).num_iters.size == 0
    with pytest.raises(Exception):
        batch.solve_batch(W.bowl(2, n=225, max_iters=3))
