"""Synthetic workloads of BASELINE.json's configs, generated identically for the CUDA path
and for the CPU checker: counter-based SplitMix64, so instance i's data does not depend on sharding.
"""
from __future__ import annotations

import math

import numpy as np

from . import _abi
from .problem import BatchProblem

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix_u01(seed: int, index: np.ndarray) -> np.ndarray:
    """u in [0,1): the (index+1)-th output of SplitMix64 seeded with `seed`, (x >> 11) * 2^-53."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (np.asarray(index, dtype=np.uint64) + np.uint64(1)) * _GOLD
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _uniform(seed, count, width, lo, hi, first=0):
    idx = (np.arange(first, first + count, dtype=np.uint64)[:, None] * np.uint64(width)
           + np.arange(width, dtype=np.uint64)[None, :])
    u = splitmix_u01(seed, idx)
    return lo + (hi - lo) * u


SEED = 0x5EED0000


def example1(count: int = 1) -> BatchProblem:
    """C1: src/example1.rs:39-50 — Ell::new_with_scalar(10,[0,0]), gamma=-inf, tol 1e-10 → 25 iterations."""
    return BatchProblem(oracle_id=_abi.ORACLE_EXAMPLE1, n=2, count=count, kappa0=10.0, xc0=np.zeros(2),
                        gamma0=-math.inf, max_iters=2000, tolerance=1e-10)


PROFIT_REF = np.array([20.0, 40.0, 30.5, 0.1, 0.4, 10.0, 35.0])  # src/oracles/profit_oracle.rs:207-223


def profit(count: int, first: int = 0) -> BatchProblem:
    """C2: ProfitOracle, Ell::new([100,100],[0,0]), gamma0=0, default Options.  Global instance 0 is the
    reference test's own (→ 83 iterations)."""
    u = _uniform(SEED + 2, count, 7, 0.0, 1.0, first)
    lo = np.array([15.0, 30.0, 20.0, 0.05, 0.3, 8.0, 30.0])
    hi = np.array([25.0, 50.0, 40.0, 0.15, 0.5, 12.0, 40.0])
    p = lo + (hi - lo) * u
    if first == 0 and count > 0:
        p[0] = PROFIT_REF
    return BatchProblem(oracle_id=_abi.ORACLE_PROFIT, n=2, count=count, kappa0=1.0, xc0=np.zeros(2),
                        diag0=np.array([100.0, 100.0]), gamma0=0.0, params=p)


def lowpass_params(count: int, first: int = 0) -> np.ndarray:
    u = _uniform(SEED + 3, count, 3, 0.0, 1.0, first)
    wpass = 0.10 + 0.04 * u[:, 0]
    dp = 0.02 + 0.01 * u[:, 1]
    ds = 0.10 + 0.05 * u[:, 2]
    if first == 0 and count > 0:
        wpass[0], dp[0], ds[0] = 0.12, 0.025, 0.125
    wstop = wpass + 0.08
    d1 = 20.0 * np.log10(1.0 + dp)
    lp = 10.0 ** (-d1 / 20.0)
    up = 10.0 ** (d1 / 20.0)
    return np.stack([wpass, wstop, lp * lp, up * up, ds * ds], axis=1)


def lowpass(count: int, table: np.ndarray, n: int = 32, first: int = 0, max_iters: int = 50000) -> BatchProblem:
    """C3: LowpassOracle::new(n, wpass, wstop, lp_sq, up_sq, sp_sq) with sane bounds, Ell::new_with_scalar(40, 0),
    gamma0 = sp_sq, Options{50000, 1e-14} (src/oracles/lowpass_oracle.rs:175-189).  `table` is the spectrum
    (ellcuda_lowpass_table)."""
    p = lowpass_params(count, first)
    return BatchProblem(oracle_id=_abi.ORACLE_LOWPASS, n=n, count=count, kappa0=40.0, xc0=np.zeros(n),
                        gamma0=p[:, 4].copy(), params=p, table=table, aux=[15 * n],
                        max_iters=max_iters, tolerance=1e-14)


def bowl(count: int, n: int = 32, first: int = 0, max_iters: int = 20000, spread: float = 1.0) -> BatchProblem:
    """C3': shifted quadratic bowl (tests/integration_test.rs:90-105), t in U[-spread,spread]^n,
    Ell::new_with_scalar(100, 0), gamma0=+inf, tol 1e-10."""
    t = _uniform(SEED + 4, count, n, -spread, spread, first)
    return BatchProblem(oracle_id=_abi.ORACLE_BOWL, n=n, count=count, kappa0=100.0, xc0=np.zeros(n),
                        gamma0=math.inf, params=t, max_iters=max_iters, tolerance=1e-10)


def lmi_table(n: int = 64, m: int = 16) -> np.ndarray:
    """shared F_k (k<n) symmetric m×m U[-1,1], then B = I + 0.0125·sym(U[-1,1])."""
    u = _uniform(SEED + 5, n + 1, m * m, -1.0, 1.0).reshape(n + 1, m, m)
    low = np.tril(u)
    sym = low + np.transpose(np.tril(u, -1), (0, 2, 1))
    f = sym[:n]
    b = np.eye(m) + 0.0125 * sym[n]
    return np.concatenate([f.reshape(-1), b.reshape(-1)])


def lmi(count: int, n: int = 64, m: int = 16, first: int = 0, general: bool = False, table=None,
        upper_scale: float = 0.1) -> BatchProblem:
    """C4: LMIOracle feasibility with EllStable.  C4a (general=False): EllStable::new_with_scalar(100, xc0);
    C4b (general=True): new_with_matrix(100, U0, xc0), U0 = unit diagonal + strict upper U[-s,s], s = upper_scale.
    (Under the reference's recurrence a non-zero upper triangle soon drives omega negative — runs end in NoSoln after
    3-40 iterations depending on s; that is the reference's behaviour and is what parity is held to.)"""
    if table is None:
        table = lmi_table(n, m)
    xc0 = _uniform(SEED + 6, count, n, -1.0, 1.0, first)
    mq0 = None
    if general:
        u = _uniform(SEED + 7, 1, n * n, -upper_scale, upper_scale).reshape(n, n)
        mq0 = np.triu(u, 1) + np.eye(n)
    return BatchProblem(oracle_id=_abi.ORACLE_LMI, n=n, count=count, mode=_abi.MODE_FEAS, kind=_abi.KIND_ELLSTABLE,
                        kappa0=100.0, xc0=xc0, mq0=mq0, table=table, aux=[m], max_iters=2000, tolerance=1e-20)


def lmi0(count: int, n: int = 64, m: int = 16, first: int = 0) -> BatchProblem:
    """LMI0Oracle (src/oracles/lmi0_oracle.rs) as a feasibility problem: find x with sum_k F_k x_k > 0.  F_0 = I +
    0.25*sym(U[-1,1]) keeps the cone non-trivial (x = e_0 is inside), the other F_k are lmi_table's; EllStable::
    new_with_scalar(100, xc0), xc0 in U[-1,1]^n."""
    tab = lmi_table(n, m)[: n * m * m].copy()
    f0 = tab[: m * m].reshape(m, m)
    tab[: m * m] = (np.eye(m) + 0.25 * f0).reshape(-1)
    xc0 = _uniform(SEED + 8, count, n, -1.0, 1.0, first)
    return BatchProblem(oracle_id=_abi.ORACLE_LMI0, n=n, count=count, mode=_abi.MODE_FEAS, kind=_abi.KIND_ELLSTABLE,
                        kappa0=100.0, xc0=xc0, table=tab, aux=[m], max_iters=2000, tolerance=1e-20)


def large_n_grad(n: int, k: int) -> np.ndarray:
    """C5: cut k of the n-dimensional single ellipsoid, g_k[j] = (2u-1)/sqrt(n)."""
    u = splitmix_u01(SEED + 8, np.uint64(k) * np.uint64(n) + np.arange(n, dtype=np.uint64))
    return (2.0 * u - 1.0) / math.sqrt(n)


LARGE_N_CUT = (0.01, 0.10)  # ParallelCut(beta0, Some(beta1)) of C5
