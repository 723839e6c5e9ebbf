"""Seeded random sweep over shapes, flags and oracle parameters: every batched kernel against the CPU checker, bit for bit.
The fixed-case tests pin known situations; this one walks the corners nobody thought of (degenerate bands, huge / tiny
kappa, targets outside the ellipsoid, zero iterations, tolerances that stop immediately)."""
import math

import numpy as np
import pytest

from ellalgo_rs_b200 import _abi as A
from ellalgo_rs_b200 import batch
from ellalgo_rs_b200 import workloads as W
from ellalgo_rs_b200.problem import BatchProblem
from parity import assert_parity

pytestmark = pytest.mark.gpu


def _check(ref, prob, what):
    assert_parity(batch.solve_batch(prob), ref.solve(prob), exact=True, what=what)


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_bowl_any_n(ref, seed):  # batch_warp_kernel<OrBowl>, n = 3..32 (generic and FIXED32 variants)
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(3, 33)) if seed else 32
    count = int(rng.integers(1, 70))
    prob = W.bowl(count, n=n, max_iters=int(rng.integers(0, 900)), spread=float(rng.choice([0.3, 1.0, 3.0])))
    prob.kappa0 = rng.uniform(1e-3, 1e3, count) if seed % 2 else np.full(count, float(rng.choice([1e-6, 1.0, 100.0, 1e8])))
    prob.tolerance = float(rng.choice([1e-10, 1e-4, 1e3]))
    prob.no_defer_trick = bool(rng.integers(0, 2))
    prob.use_parallel_cut = bool(rng.integers(0, 2))
    if seed % 3 == 0:
        prob.kappa0 = np.full(count, 1.0)
        prob.diag0 = rng.uniform(0.5, 200.0, (count, n))
    _check(ref, prob, f"bowl fuzz seed={seed} n={n} count={count}")


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_lowpass(ref, seed):  # lowpass32_tmem_kernel (n = 32) and batch_warp_kernel<OrLowpass> (n < 32)
    rng = np.random.default_rng(2000 + seed)
    n = 32 if seed < 4 else int(rng.choice([8, 12, 20, 31]))
    t = batch.lowpass_table(n)
    count = int(rng.integers(1, 40))
    prob = W.lowpass(count, t, n=n, max_iters=int(rng.integers(1, 1200)))
    p = prob.params
    p[:, 0] = rng.uniform(0.0, 0.6, count)                       # wpass
    p[:, 1] = p[:, 0] + rng.uniform(-0.05, 0.4, count)           # wstop, sometimes below wpass (empty transition band)
    p[:, 1] = np.clip(p[:, 1], 0.0, 0.999)
    dp = rng.uniform(0.005, 0.2, count)
    p[:, 2] = (1.0 - dp) ** 2
    p[:, 3] = (1.0 + dp) ** 2
    if seed == 1:
        p[::2, 2], p[::2, 3] = p[::2, 3].copy(), p[::2, 2].copy()   # lp_sq > up_sq: the reference's own degenerate case
    p[:, 4] = rng.uniform(1e-4, 0.5, count)
    prob.gamma0 = p[:, 4].copy()
    prob.kappa0 = rng.uniform(0.5, 200.0, count)
    prob.use_parallel_cut = bool(rng.integers(0, 2))
    prob.no_defer_trick = bool(rng.integers(0, 2))
    prob.tolerance = float(rng.choice([1e-14, 1e-6]))
    _check(ref, prob, f"lowpass fuzz seed={seed} n={n} count={count}")


@pytest.mark.parametrize("seed", range(8))
def test_fuzz_cta_kernel(ref, seed):  # batch_ellcta_kernel: bowl n = 33..224, lowpass n = 48..224 (packed triangle in shared memory)
    rng = np.random.default_rng(2500 + seed)
    if seed < 4:
        n = int(rng.integers(33, 225))
        count = int(rng.integers(1, 30))
        prob = W.bowl(count, n=n, max_iters=int(rng.integers(0, 500)), spread=float(rng.choice([0.3, 1.0, 3.0])))
        prob.kappa0 = rng.uniform(1e-3, 1e3, count) if seed % 2 else np.full(count, float(rng.choice([1e-6, 1.0, 100.0, 1e8])))
        prob.tolerance = float(rng.choice([1e-10, 1e-4, 1e3]))
        if seed == 2:
            prob.kappa0 = np.full(count, 1.0)
            prob.diag0 = rng.uniform(0.5, 200.0, (count, n))
        if seed == 3:
            prob.kappa0 = np.full(count, 1.0)
            prob.mq0 = np.ascontiguousarray(50.0 * (np.eye(n)[None] + rng.uniform(-1, 1, (count, n, n)) / (2 * n)))   # non-symmetric
    else:
        n = int(rng.integers(48, 225))
        count = int(rng.integers(1, 8))
        prob = W.lowpass(count, batch.lowpass_table(n), n=n, max_iters=int(rng.integers(1, 500)))
        p = prob.params
        p[:, 0] = rng.uniform(0.0, 0.6, count)
        p[:, 1] = np.clip(p[:, 0] + rng.uniform(-0.05, 0.4, count), 0.0, 0.999)   # sometimes below wpass: empty transition band
        dp = rng.uniform(0.005, 0.2, count)
        p[:, 2] = (1.0 - dp) ** 2
        p[:, 3] = (1.0 + dp) ** 2
        if seed == 5:
            p[::2, 2], p[::2, 3] = p[::2, 3].copy(), p[::2, 2].copy()              # lp_sq > up_sq
        p[:, 4] = rng.uniform(1e-4, 0.5, count)
        prob.gamma0 = p[:, 4].copy()
        prob.kappa0 = rng.uniform(0.5, 200.0, count)
        prob.tolerance = float(rng.choice([1e-14, 1e-6]))
    prob.use_parallel_cut = bool(rng.integers(0, 2))
    prob.no_defer_trick = bool(rng.integers(0, 2))
    _check(ref, prob, f"cta fuzz seed={seed} n={n} count={count}")


@pytest.mark.parametrize("seed", range(5))
def test_fuzz_lmi_family(ref, seed):  # batch_stable_kernel: LMI / LMIOld / LMI0, diagonal and general starts
    rng = np.random.default_rng(3000 + seed)
    n = int(rng.integers(2, 65))
    m = int(rng.integers(1, 21))
    count = int(rng.integers(1, 50))
    which = [A.ORACLE_LMI, A.ORACLE_LMI_OLD, A.ORACLE_LMI0][seed % 3]
    prob = W.lmi0(count, n=n, m=m) if which == A.ORACLE_LMI0 else W.lmi(count, n=n, m=m)
    prob.oracle_id = which
    prob.max_iters = int(rng.integers(0, 400))
    prob.kappa0 = rng.uniform(0.1, 500.0, count)
    if seed >= 3:
        u = rng.uniform(-1e-3, 1e-3, (n, n))
        prob.mq0 = np.triu(u, 1) + np.diag(rng.uniform(0.5, 2.0, n))
    elif seed == 2:
        prob.kappa0 = np.full(count, 1.0)
        prob.diag0 = rng.uniform(1.0, 300.0, (count, n))
    _check(ref, prob, f"lmi-family fuzz seed={seed} oracle={which} n={n} m={m} count={count}")


@pytest.mark.parametrize("seed", range(4))
def test_fuzz_small_oracles(ref, seed):  # batch_small_kernel: profit family, both spaces, odd options
    rng = np.random.default_rng(4000 + seed)
    count = int(rng.integers(1, 3000))
    prob = W.profit(count)
    prob.kind = A.KIND_ELLSTABLE if seed % 2 else A.KIND_ELL
    prob.max_iters = int(rng.integers(0, 300))
    prob.tolerance = float(rng.choice([1e-20, 1e-8, 1e-2]))
    prob.use_parallel_cut = bool(rng.integers(0, 2))
    if seed == 2:
        prob.oracle_id, prob.mode = A.ORACLE_PROFIT_Q, A.MODE_OPTIM_Q
    if seed == 3:
        prob.diag0 = rng.uniform(1.0, 400.0, (count, 2))
        prob.xc0 = rng.uniform(-2.0, 6.0, (count, 2))
    exact = True   # exp/ln on the device are glibc's (csrc/glibc_math.cuh)
    assert_parity(batch.solve_batch(prob), ref.solve(prob), exact=exact, what=f"profit fuzz seed={seed} count={count}")


@pytest.mark.parametrize("seed", range(8))
def test_fuzz_single_ellipsoid_trajectory(ref, seed):
    """ellcuda_update on a handle (sweep + finish kernels, deferred rank-1): random n (tile-unaligned), random start, a random
    walk over update kinds and cut depths that mixes Success / NoSoln / NoEffect, a clone and a set_xc in the middle —
    status, tsq, kappa, xc after every step and the full matrix at the end, bit for bit"""
    from ellalgo_rs_b200.cutting_plane import ParallelCut, SingleCut
    from ellalgo_rs_b200.ell import Ell

    rng = np.random.default_rng(5000 + seed)
    n = int(rng.choice([1, 2, 3, 31, 63, 64, 65, 127, 129, 200, 333, 511, 700])) if seed else 1
    xc = rng.uniform(-1, 1, n)
    no_defer = bool(rng.integers(0, 2))
    use_par = bool(rng.integers(0, 2))
    kw = {}
    if seed % 3 == 1:
        kw["diag"] = rng.uniform(0.5, 2.0, n)
    elif seed % 3 == 2:
        a = rng.uniform(-1, 1, (n, n)) / n
        kw["mq"] = np.eye(n) + a + a.T
    kappa = 1.0 if kw else float(rng.uniform(0.5, 5.0))
    e = Ell(kappa, xc, no_defer_trick=no_defer, use_parallel_cut=use_par, **kw)
    s = ref.Space(A.KIND_ELL, n, kappa=kappa, xc=xc, no_defer_trick=no_defer, use_parallel_cut=use_par, **kw)
    for k in range(14):
        g = rng.uniform(-1, 1, n) / math.sqrt(n)
        tau = math.sqrt(max(s.state()["tsq"], 1e-3)) if k else 1.0
        which = int(rng.integers(0, 3))
        b0 = float(rng.choice([-0.5, 0.0, 0.05, 0.3, 0.9, 1.5]) * tau)
        if rng.integers(0, 2):
            cut = SingleCut(b0)
            st_r = s.update(which, g, b0)
        else:
            b1 = None if rng.integers(0, 4) == 0 else float(b0 + rng.choice([-0.1, 0.05, 0.4, 2.0]) * tau)
            cut = ParallelCut(b0, b1)
            st_r = s.update(which, g, b0, b1, single=False)
        st_g = [e.update_bias_cut, e.update_central_cut, e.update_q][which]((g, cut))
        r = s.state()
        assert int(st_g) == st_r, f"seed {seed} n {n} step {k}: status {int(st_g)} vs {st_r}"
        same = lambda a, b: np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)   # n = 1: cst1 = 1/0, NaNs as in the reference
        assert same(e.tsq(), r["tsq"]) and same(e.kappa, r["kappa"]), f"seed {seed} n {n} step {k}: tsq/kappa"
        assert same(e.xc(), r["xc"]), f"seed {seed} n {n} step {k}: xc"
        if k == 6:
            e = e.clone()                      # the pending rank-1 must travel with the copy
            x2 = rng.uniform(-1, 1, n)
            e.set_xc(x2)
            s.set_xc(x2)
    assert np.array_equal(e.mq, s.state()["mq"], equal_nan=True), f"seed {seed} n {n}: final matrix"
