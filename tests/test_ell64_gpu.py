"""GPU parity tests of batched `Ell` for 32 < n <= 64: the quadratic bowl and LowpassOracle under cutting_plane_optim
(routed to the CTA-per-problem kernel csrc/batch_ellcta.cu, except lowpass below n = 48, and — in a child process with
ELLCUDA_CTA_MIN_N=65 — through the tensor-memory kernel csrc/batch_ell64.cu), and LMIOracle / LMIOldOracle / LMI0Oracle on
the `Ell` space under cutting_plane_feas with M in tensor memory — the reference runs its LMI oracles under both search
spaces (tests/lmi_tests.rs:174-225).  Bit for bit against the CPU checker."""
import os
import subprocess
import sys

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


@pytest.mark.parametrize("n,count,spread,max_iters", [(33, 40, 1.0, 500), (48, 65, 1.0, 700), (64, 130, 1.0, 900), (64, 40, 3.0, 400),
                                                      (50, 1201, 1.0, 60)])
def test_bowl_n_up_to_64(ref, n, count, spread, max_iters):  # tests/integration_test.rs:90-105 at SURVEY K1's "n = 64"
    prob = W.bowl(count, n=n, spread=spread, max_iters=max_iters)
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what=f"bowl n={n}")
    assert g.num_iters.max() > 20


def test_bowl64_runs_to_tolerance(ref):
    prob = W.bowl(3, n=40, max_iters=20000)
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what="bowl n=40 to tolerance")
    assert (g.exit_status == A.EXIT_TOL).all() and g.num_iters.min() > 5000


@pytest.mark.parametrize("n", [36, 64])
def test_bowl64_flags_and_starts(ref, n):
    rng = np.random.default_rng(n)
    for kw in (dict(no_defer_trick=True), dict(use_parallel_cut=False), dict(max_iters=1), dict(max_iters=0)):
        prob = W.bowl(21, n=n, max_iters=kw.get("max_iters", 150))
        prob.no_defer_trick = kw.get("no_defer_trick", False)
        prob.use_parallel_cut = kw.get("use_parallel_cut", True)
        g, r = both(ref, prob)
        assert_parity(g, r, exact=True, what=f"bowl n={n} {kw}")
    prob = W.bowl(19, n=n, max_iters=200)                      # Ell::new(val, xc): per-instance diagonals
    prob.kappa0 = np.ones(1)
    prob.diag0 = rng.uniform(50.0, 150.0, (19, n))
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what=f"bowl n={n} diag start")
    a = rng.uniform(-1, 1, (19, n, n)) / n                     # Ell::new_with_matrix, symmetric dense
    prob.diag0 = None
    prob.mq0 = np.ascontiguousarray(100.0 * (np.eye(n) + a + np.transpose(a, (0, 2, 1))))
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what=f"bowl n={n} matrix start")
    prob.mq0 = np.ascontiguousarray(100.0 * (np.eye(n) + a))   # ... and non-symmetric (src/ell.rs:31-41, 117-128)
    prob.no_defer_trick = True
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what=f"bowl n={n} non-symmetric matrix start")


@pytest.mark.parametrize("n,count,max_iters", [(40, 9, 1500), (64, 7, 1200), (33, 5, 800)])
def test_lowpass_n_up_to_64(ref, n, count, max_iters):  # src/oracles/lowpass_oracle.rs with n > 32 taps (stress_tests.rs uses 128)
    t = batch.lowpass_table(n)
    prob = W.lowpass(count, t, n=n, max_iters=max_iters)
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what=f"lowpass n={n}")
    assert g.num_iters.max() == max_iters or (g.exit_status != A.EXIT_MAXIT).any()
    assert g.has_xbest.any() or max_iters < 1500


def test_lowpass64_literal_case(ref):  # create_lowpass_case(n): degenerate bounds -> NoSoln at iteration 0
    n = 48
    t = batch.lowpass_table(n)
    p = ref.lowpass_case_params()
    prob = BatchProblem(oracle_id=A.ORACLE_LOWPASS, n=n, count=2, kappa0=40.0, xc0=np.zeros(n), gamma0=p[4], params=p, table=t,
                        aux=[15 * n], max_iters=50000, tolerance=1e-14)
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what="lowpass literal n=48")


@pytest.mark.parametrize("n,m,count", [(64, 16, 96), (20, 7, 40), (33, 16, 17), (64, 5, 33), (3, 2, 9)])
@pytest.mark.parametrize("oracle", [A.ORACLE_LMI, A.ORACLE_LMI_OLD])
def test_lmi_on_ell(ref, n, m, count, oracle):  # configs[3] with the Ell space (tests/lmi_tests.rs:174-206 run both spaces)
    prob = W.lmi(count, n=n, m=m)
    prob.kind = A.KIND_ELL
    prob.oracle_id = oracle
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what=f"lmi on Ell n={n} m={m}")
    assert (g.exit_status == A.EXIT_FEASIBLE).any() and g.num_iters.max() > 3


def test_lmi0_on_ell(ref):
    prob = W.lmi0(40, n=64, m=16)
    prob.kind = A.KIND_ELL
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what="lmi0 on Ell")
    assert (g.num_iters > 0).any()


def test_lmi_on_ell_matrix_start_and_flags(ref):
    n, m, cnt = 40, 9, 23
    rng = np.random.default_rng(4)
    prob = W.lmi(cnt, n=n, m=m)
    prob.kind = A.KIND_ELL
    a = rng.uniform(-1, 1, (n, n)) / n
    prob.mq0 = np.ascontiguousarray(np.eye(n) + a)             # non-symmetric new_with_matrix start
    for nd in (False, True):
        prob.no_defer_trick = nd
        g, r = both(ref, prob)
        assert_parity(g, r, exact=True, what=f"lmi on Ell, matrix start, no_defer={nd}")


def test_ell64_more_problems_than_resident_warps_and_empty(ref):
    prob = W.bowl(148 * 8 * 2 + 5, n=64, max_iters=25)
    g, r = both(ref, prob)
    assert_parity(g, r, exact=True, what="bowl n=64, 2 waves")
    assert batch.solve_batch(W.bowl(0, n=64)).num_iters.size == 0


def test_tmem_kernel_bowl_and_lowpass_when_selected():
    """the tensor-memory kernel's own bowl / lowpass paths (the default routing prefers the CTA kernel for them)"""
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, ELLCUDA_CTA_MIN_N="65")
    out = subprocess.run([sys.executable, os.path.join(here, "ell64_tmem_check.py")], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "tmem kernel ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
