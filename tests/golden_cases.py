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
