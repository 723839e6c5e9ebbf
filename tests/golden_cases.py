"""The cases frozen in tests/golden/golden_v1.npz (inputs are regenerated from the seeded workload generators)."""
import math

import numpy as np

from ellalgo_rs_b200 import _abi as A
from ellalgo_rs_b200 import workloads as W
from ellalgo_rs_b200.problem import BatchProblem


def _lowpass(R):
    return W.lowpass(3, R.lowpass_table(32) if hasattr(R, "lowpass_table") else R, max_iters=50000)


def _lowpass_bands(R):
    """awkward band layouts (short bands, empty transition band, edges on multiples of 32): the TMEM kernel's rounds"""
    p = W.lowpass(5, R.lowpass_table(32) if hasattr(R, "lowpass_table") else R, max_iters=1500)
    for i, (wp, ws) in enumerate([(0.001, 0.002), (0.12, 0.12), (31.5 / 479, 63.5 / 479), (0.05, 0.995), (0.90, 0.9979)]):
        p.params[i, 0], p.params[i, 1] = wp, ws
    return p


CASES = {
    "example1": lambda R: W.example1(),
    "profit": lambda R: W.profit(64),
    "profit_q": lambda R: (lambda p: (setattr(p, "oracle_id", A.ORACLE_PROFIT_Q), setattr(p, "mode", A.MODE_OPTIM_Q), p)[2])(W.profit(16)),
    "quasicvx_stable": lambda R: BatchProblem(oracle_id=A.ORACLE_QUASICVX, n=2, count=1, kind=A.KIND_ELLSTABLE, kappa0=1.0,
                                              diag0=[10.0, 10.0], gamma0=0.0, tolerance=1e-8),
    "bowl32": lambda R: W.bowl(8, n=32, max_iters=20000),
    "bowl5": lambda R: BatchProblem(oracle_id=A.ORACLE_BOWL, n=5, count=1, kappa0=10.0, gamma0=math.inf,
                                    params=[1.0, 2, 3, 4, 5], max_iters=2000, tolerance=1e-10),
    "lowpass32": _lowpass,
    "lmi_diag": lambda R: W.lmi(6, n=64, m=16),
    "lmi_general": lambda R: W.lmi(6, n=64, m=16, general=True, upper_scale=1e-6),
    "lmi_old": lambda R: (lambda p: (setattr(p, "oracle_id", A.ORACLE_LMI_OLD), p)[1])(W.lmi(4, n=20, m=7)),
    "lmi0": lambda R: W.lmi0(6, n=64, m=16),
    "lowpass32_bands": lambda R: _lowpass_bands(R),
}


def run_stepwise(make_space, update, state):
    """4 updates of an n=48 Ell from a dense symmetric start with C5's parallel cut; returns the final state."""
    n = 48
    rng = np.random.default_rng(123)
    a = rng.uniform(-1, 1, (n, n)) / n
    s = make_space(n=n, kappa=1.5, mq=np.eye(n) + a + a.T, xc=rng.uniform(-1, 1, n))
    for k in range(4):
        update(s, 0, W.large_n_grad(n, k), *W.LARGE_N_CUT)
    st = state(s)
    return {"xc": st["xc"], "mq": st["mq"], "kappa": np.array(st["kappa"]), "tsq": np.array(st["tsq"])}


# ---------------------------------------------------------------------------------- round 2: tests/golden/golden_v2.npz
def _asym_batch(n, count, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(scale * (np.eye(n)[None] + rng.uniform(-1, 1, (count, n, n)) / (2 * n)))


def _with(p, **kw):
    for k, v in kw.items():
        setattr(p, k, v)
    return p


CASES_V2 = {
    "bowl48": lambda R: W.bowl(6, n=48, max_iters=3000),                                   # M in tensor memory (batch_ell64.cu)
    "bowl64_nonsym": lambda R: _with(W.bowl(5, n=64, max_iters=600), mq0=_asym_batch(64, 5, 11, 100.0), kappa0=np.ones(1)),
    "bowl32_nonsym": lambda R: _with(W.bowl(5, n=32, max_iters=600), mq0=_asym_batch(32, 5, 12, 100.0), kappa0=np.ones(1)),
    "lowpass40": lambda R: W.lowpass(3, R.lowpass_table(40), n=40, max_iters=2500),
    "lmi_on_ell": lambda R: _with(W.lmi(6, n=64, m=16), kind=A.KIND_ELL),                 # tests/lmi_tests.rs:174-206 run both spaces
    "lmi0_on_ell": lambda R: _with(W.lmi0(5, n=64, m=16), kind=A.KIND_ELL),
    "profit_nonsym": lambda R: _with(W.profit(8), diag0=None, mq0=np.ascontiguousarray(
        np.array([[100.0, 1.5], [-0.75, 100.0]])[None] * np.ones((8, 1, 1)))),
}


def run_stepwise_nonsym(make_space, update, state):
    """an n=40 Ell from a NON-symmetric matrix: two failing cuts first (rows multiplied as given, nothing mirrored), then four
    successful ones (the first rewrites the upper triangle from the lower one, src/ell.rs:117-128); returns the final state."""
    n = 40
    rng = np.random.default_rng(321)
    mq = np.eye(n) + rng.uniform(-1, 1, (n, n)) / (2 * n)
    s = make_space(n=n, kappa=1.25, mq=mq, xc=rng.uniform(-1, 1, n))
    update(s, 0, W.large_n_grad(n, 0), 5.0, 4.0)          # beta1 < beta0: NoSoln
    update(s, 0, W.large_n_grad(n, 1), 1e6, 2e6)          # tsq < beta^2: NoSoln
    mid = state(s)
    for k in range(2, 6):
        update(s, 0, W.large_n_grad(n, k), *W.LARGE_N_CUT)
    st = state(s)
    return {"xc": st["xc"], "mq": st["mq"], "kappa": np.array(st["kappa"]), "tsq": np.array(st["tsq"]), "mq_before_first_success": mid["mq"]}


# ---------------------------------------------------------------------------------- round 2: tests/golden/golden_v3.npz
# batched Ell above n = 64: one CTA per problem, packed lower triangle of M in shared memory (batch_ellcta.cu)
CASES_V3 = {
    "bowl128": lambda R: W.bowl(4, n=128, max_iters=2500),
    "bowl100_nonsym": lambda R: _with(W.bowl(4, n=100, max_iters=500), mq0=_asym_batch(100, 4, 13, 100.0), kappa0=np.ones(1)),
    "bowl224_no_defer": lambda R: _with(W.bowl(2, n=224, max_iters=300), no_defer_trick=True),
    "lowpass128": lambda R: W.lowpass(3, R.lowpass_table(128), n=128, max_iters=2000),    # tests/stress_tests.rs:6-14's dimension
    "lowpass72": lambda R: W.lowpass(3, R.lowpass_table(72), n=72, max_iters=2000),
}
