"""The reference's own integration suites, re-run with the unchanged driver logic over (a) the CPU checker and (b) the GPU
search space (`ellalgo_rs_b200.ell.Ell`, i.e. libellcuda.so behind the SearchSpace surface), with the reference's asserts:

    tests/regression_tests.rs        tests/numerical_stability.rs     tests/integration_test.rs
    tests/quasicvx2_tests.rs         tests/stress_tests.rs            tests/cutting_plane_tests.rs:328-379 (BSearchAdaptor)

The oracles are user code in the reference (closures over `Arr`), so here they are host-side Python objects — the drop-in
story of INTEGRATION.md §1: user oracle on the host, ellipsoid on the device.  Every arithmetic expression keeps the
reference's association (`powi(2)` is x*x; Python floats are IEEE binary64 without contraction; exp/sqrt are libm's).
The CPU variants pin the checker against these suites; the GPU variants must in addition equal the checker's run bit for bit.
"""
import math

import numpy as np
import pytest

from ellalgo_rs_b200 import _abi as A
from ellalgo_rs_b200.cutting_plane import (BSearchAdaptor, CutStatus, Options, ParallelCut, SingleCut, bsearch,
                                           cutting_plane_optim)

INF = math.inf
SPACES = ["checker", pytest.param("gpu", marks=pytest.mark.gpu)]


class CheckerSpace:
    """SearchSpace surface (src/cutting_plane.rs:154-182) over the CPU checker's Ell"""

    def __init__(self, ref, kappa, xc, _space=None):
        self.s = _space if _space is not None else ref.Space(A.KIND_ELL, len(xc), kappa=kappa, xc=np.asarray(xc, dtype=np.float64))

    def clone(self):
        return CheckerSpace(None, None, None, _space=self.s.clone())

    def set_xc(self, x):
        self.s.set_xc(np.asarray(x, dtype=np.float64))

    def xc(self):
        return self.s.state()["xc"].copy()

    def tsq(self):
        return float(self.s.state()["tsq"])

    def _update(self, which, cut):
        g, c = cut
        if isinstance(c, SingleCut):
            return CutStatus(self.s.update(which, g, c.beta))
        return CutStatus(self.s.update(which, g, c.beta0, c.beta1, single=False))

    def update_bias_cut(self, cut):
        return self._update(A.UPDATE_BIAS, cut)

    def update_central_cut(self, cut):
        return self._update(A.UPDATE_CENTRAL, cut)

    def update_q(self, cut):
        return self._update(A.UPDATE_Q, cut)


def make_space(kind, ref, kappa, xc):
    if kind == "checker":
        return CheckerSpace(ref, kappa, xc)
    from ellalgo_rs_b200.ell import Ell
    return Ell.new_with_scalar(kappa, np.asarray(xc, dtype=np.float64))


def solve(kind, ref, oracle_factory, kappa, xc, gamma, options):
    """cutting_plane_optim on the requested space; on the GPU also on the checker, and the two must agree in every bit"""
    x, niter, g = cutting_plane_optim(oracle_factory(), make_space(kind, ref, kappa, xc), gamma, options)
    if kind == "gpu":
        x2, niter2, g2 = cutting_plane_optim(oracle_factory(), CheckerSpace(ref, kappa, xc), gamma, options)
        assert niter == niter2 and (g == g2 or (g != g and g2 != g2)), (niter, niter2, g, g2)
        assert (x is None) == (x2 is None) and (x is None or np.array_equal(x, x2))
    return x, niter, g


# ------------------------------------------------------------------------------------------------ oracles (user code)
class Quadratic:  # regression_tests.rs:7-22 and its copies; integration_test.rs:7-21
    def __init__(self, a0=1.0, a1=1.0, plain=True):
        self.a0, self.a1, self.plain = a0, a1, plain

    def assess_optim(self, xc, gamma):
        if self.plain:
            grad = np.array([2.0 * xc[0], 2.0 * xc[1]])
            f = xc[0] * xc[0] + xc[1] * xc[1]
        else:   # numerical_stability.rs:14-18 / :97-100: 2.0 * a * x and a * x^2, left to right
            grad = np.array([2.0 * self.a0 * xc[0], 2.0 * self.a1 * xc[1]])
            f = self.a0 * (xc[0] * xc[0]) + self.a1 * (xc[1] * xc[1])
        if f < gamma:
            return (grad, SingleCut(f)), True, f
        return (grad, SingleCut(f)), False, gamma


class ExtremeScale:  # numerical_stability.rs:94-108: gradient (2 s x0, 2 x1), f = s x0^2 + x1^2
    def __init__(self, s):
        self.s = s

    def assess_optim(self, xc, gamma):
        grad = np.array([2.0 * self.s * xc[0], 2.0 * xc[1]])
        f = self.s * (xc[0] * xc[0]) + xc[1] * xc[1]
        if f < gamma:
            return (grad, SingleCut(f)), True, f
        return (grad, SingleCut(f)), False, gamma


class Scaling:  # regression_tests.rs:173-193: f accumulated from 0.0 over i
    def __init__(self, ndim):
        self.ndim = ndim

    def assess_optim(self, xc, gamma):
        grad = np.zeros(self.ndim)
        f = 0.0
        for i in range(self.ndim):
            grad[i] = 2.0 * xc[i]
            f += xc[i] * xc[i]
        if f < gamma:
            return (grad, SingleCut(f)), True, f
        return (grad, SingleCut(f)), False, gamma


class ShiftedBowl:  # integration_test.rs:44-60, 90-105: diff = xc - target; 2*diff; diff.dot(diff) sequential
    def __init__(self, target):
        self.t = np.asarray(target, dtype=np.float64)

    def assess_optim(self, xc, gamma):
        diff = xc - self.t
        grad = 2.0 * diff
        f = 0.0
        for d in diff:
            f += d * d
        if f < gamma:
            return (grad, SingleCut(f)), True, f
        return (grad, SingleCut(f)), False, gamma


class MyQuasicvx:  # quasicvx2_tests.rs:14-71
    def __init__(self):
        self.idx = -1

    def assess_optim(self, xc, gamma):
        x, y = float(xc[0]), float(xc[1])
        for _ in range(3):
            self.idx += 1
            if self.idx == 3:
                self.idx = 0
            if self.idx == 0:
                tmp = math.exp(x)
                grad, fj = np.array([tmp, -1.0]), tmp - y
            elif self.idx == 1:
                grad, fj = np.array([0.0, -1.0]), -y
            else:
                grad, fj = np.array([-1.0, 0.0]), -x
            if fj > 0.0:
                return (grad, SingleCut(fj)), False, gamma
        tmp2 = math.sqrt(x)
        fj = -tmp2 + gamma * y
        if fj > 0.0:
            return (np.array([-0.5 / tmp2, gamma]), SingleCut(fj)), False, gamma
        gamma = tmp2 / y
        return (np.array([-0.5 / tmp2, gamma]), SingleCut(0.0)), True, gamma


class LowpassHost:  # src/oracles/lowpass_oracle.rs:7-150 as a host oracle (any n; the batched device oracle covers n <= 32)
    def __init__(self, ref, n, params):
        self.n, self.mdim = n, 15 * n
        self.spectrum = ref.lowpass_table(n).reshape(self.mdim, n)
        wpass, wstop, self.lp_sq, self.up_sq, self.sp_sq = (float(v) for v in params)
        self.nwpass = int(math.floor(wpass * (self.mdim - 1))) + 1
        self.nwstop = int(math.floor(wstop * (self.mdim - 1))) + 1
        self.idx1, self.idx2, self.idx3 = -1, self.nwpass - 1, self.nwstop - 1
        self.fmax, self.kmax = -INF, -1

    def _dot(self, k, x):
        s = 0.0
        for a, b in zip(self.spectrum[k], x):
            s += a * b
        return s

    def _assess_feas(self, x):
        for _ in range(self.nwpass):
            self.idx1 += 1
            if self.idx1 == self.nwpass:
                self.idx1 = 0
            v = self._dot(self.idx1, x)
            if v > self.up_sq:
                return self.spectrum[self.idx1].copy(), ParallelCut(v - self.up_sq, v - self.lp_sq)
            if v < self.lp_sq:
                return -self.spectrum[self.idx1], ParallelCut(-v + self.lp_sq, -v + self.up_sq)
        self.fmax, self.kmax = -INF, -1
        for _ in range(self.nwstop, self.mdim):
            self.idx3 += 1
            if self.idx3 == self.mdim:
                self.idx3 = self.nwstop
            v = self._dot(self.idx3, x)
            if v > self.sp_sq:
                return self.spectrum[self.idx3].copy(), ParallelCut(v - self.sp_sq, v)
            if v < 0.0:
                return -self.spectrum[self.idx3], ParallelCut(-v, -v + self.sp_sq)
            if v > self.fmax:
                self.fmax, self.kmax = v, self.idx3
        for _ in range(self.nwpass, self.nwstop):
            self.idx2 += 1
            if self.idx2 == self.nwstop:
                self.idx2 = self.nwpass
            v = self._dot(self.idx2, x)
            if v < 0.0:
                return -self.spectrum[self.idx2], ParallelCut(-v, None)
        if x[0] < 0.0:
            g = np.zeros(self.n)
            g[0] = -1.0
            return g, ParallelCut(-x[0], None)
        return None

    def assess_optim(self, x, gamma):
        self.sp_sq = gamma
        cut = self._assess_feas(x)
        if cut is not None:
            return cut, False, gamma
        return (self.spectrum[self.kmax].copy(), ParallelCut(0.0, self.fmax)), True, self.fmax


# ------------------------------------------------------------------------------------------------ regression_tests.rs
@pytest.mark.parametrize("kind", SPACES)
def test_regression_suite(ref, kind):
    x, niter, gamma = solve(kind, ref, Quadratic, 10.0, [3.0, 3.0], INF, Options(1000, 1e-10))       # :6-41
    assert niter < 1000 and gamma < 10.0
    for start in ([3.0, 3.0], [-5.0, 5.0], [10.0, -10.0]):                                            # :43-90
        x, _, _ = solve(kind, ref, Quadratic, 20.0, start, INF, Options(2000, 1e-10))
        assert x is not None and abs(x[0]) < 10.0 and abs(x[1]) < 10.0
    x, niter, gamma = solve(kind, ref, Quadratic, 10.0, [3.0, 3.0], INF, Options(500, 1e-12))        # :92-169
    assert niter > 0 and math.isfinite(gamma) and 0.0 <= gamma < 100.0 and x is not None
    for ndim in (2, 4, 8):                                                                             # :171-213
        x, niter, _ = solve(kind, ref, lambda: Scaling(ndim), 10.0, [3.0] * ndim, INF, Options(3000, 1e-10))
        assert x is not None and niter < 3000
    a = solve(kind, ref, Quadratic, 10.0, [3.0, 3.0], INF, Options(500, 1e-10))                       # :215-254
    b = solve(kind, ref, Quadratic, 10.0, [3.0, 3.0], INF, Options(500, 1e-10))
    assert a[1] == b[1] and abs(a[2] - b[2]) < 1e-10 and np.abs(a[0] - b[0]).max() < 1e-10


# ------------------------------------------------------------------------------------------------ numerical_stability.rs
@pytest.mark.parametrize("kind", SPACES)
def test_numerical_stability_suite(ref, kind):
    for cond in (1e3, 1e5, 1e7):                                                                       # :5-49
        eps = 1.0 / cond
        x, _, gamma = solve(kind, ref, lambda: Quadratic(1.0 + eps, 1.0 / eps, plain=False), 10.0, [1.0, 1.0], INF,
                            Options(2000, 1e-12))
        assert x is not None and math.isfinite(gamma)
    x, _, gamma = solve(kind, ref, Quadratic, 10.0, [3.0, 3.0], INF, Options(1000, 1e-10))            # :51-86
    assert x is not None and gamma < 10.0
    for scale in (1e-6, 1e6):                                                                          # :88-122
        x, _, _ = solve(kind, ref, lambda: ExtremeScale(scale), 10.0 * math.sqrt(abs(scale)), [scale, scale], INF,
                        Options(2000, 1e-10))
        assert x is not None
    for tol in (1e-6, 1e-10, 1e-14):                                                                   # :124-158
        x, _, _ = solve(kind, ref, Quadratic, 10.0, [3.0, 3.0], INF, Options(2000, tol))
        assert x is not None
    x, _, gamma = solve(kind, ref, Quadratic, 10.0, [1000.0, -1000.0], INF, Options(3000, 1e-12))    # :160-193
    assert x is not None and gamma < 1000.0 * 1000.0 + 1000.0 * 1000.0


# ------------------------------------------------------------------------------------------------ integration_test.rs
@pytest.mark.parametrize("kind", SPACES)
def test_integration_suite(ref, kind):
    x, _, gamma = solve(kind, ref, Quadratic, 10.0, [5.0, 5.0], INF, Options(1000, 1e-10))            # :5-38
    assert x is not None and abs(x[0]) < 0.5 and abs(x[1]) < 0.5 and gamma < 1.0
    target = [2.0, -1.5]                                                                               # :40-83
    x, _, _ = solve(kind, ref, lambda: ShiftedBowl(target), 20.0, [10.0, 10.0], INF, Options(1000, 1e-10))
    assert x is not None and math.hypot(x[0] - target[0], x[1] - target[1]) < 15.0
    target = [1.0, 2.0, 3.0, 4.0, 5.0]                                                                 # :85-131
    x, _, _ = solve(kind, ref, lambda: ShiftedBowl(target), 10.0, [0.0] * 5, INF, Options(2000, 1e-10))
    assert x is not None and math.sqrt(sum((x[i] - target[i]) ** 2 for i in range(5)) / 5) < 3.0
    x, _, gamma = solve(kind, ref, Quadratic, 10.0, [3.0, 3.0], INF, Options(500, 1e-12))             # :133-167
    assert x is not None and math.isfinite(gamma) and 0.0 <= gamma < 10.0
    for start in ([1.0, 1.0], [-1.0, -1.0], [5.0, -5.0], [10.0, 0.0]):                                # :169-220
        x, _, gamma = solve(kind, ref, Quadratic, 10.0, start, INF, Options(1000, 1e-10))
        assert x is not None and gamma < 100.0


# ------------------------------------------------------------------------------------------------ quasicvx2_tests.rs
@pytest.mark.parametrize("kind", SPACES)
def test_quasicvx2_suite(ref, kind):
    x, _, _ = solve(kind, ref, MyQuasicvx, 10.0, [1.0, 1.0], 0.0, Options())                           # :73-81
    assert x is not None
    x, _, _ = solve(kind, ref, MyQuasicvx, 10.0, [100.0, 100.0], 0.0, Options())                       # :83-91
    assert x is None
    x, _, _ = solve(kind, ref, MyQuasicvx, 10.0, [1.0, 1.0], 100.0, Options())                         # :93-100
    assert x is None


# ------------------------------------------------------------------------------------------------ stress_tests.rs
@pytest.mark.parametrize("kind", SPACES)
def test_stress_suite(ref, kind):
    """create_lowpass_case(n) is degenerate (lp_sq > up_sq): both runs must simply complete, as in the
    reference; n = 128 goes through the large-n handle (host-driven here; the same case as a device batch is tests/test_ellcta_gpu.py::test_stress_lowpass_high_dimension_batched)."""
    params = ref.lowpass_case_params()
    x, niter, _ = solve(kind, ref, lambda: LowpassHost(ref, 128, params), 1.0, [0.0] * 128, float(params[4]),  # :6-14
                        Options(50000, 1e-14))
    assert x is None and niter == 0
    x, niter, _ = solve(kind, ref, lambda: LowpassHost(ref, 32, params), 1.0, [0.0] * 32, 1e-12, Options(50000, 1e-14))  # :16-24
    assert x is None and niter == 0
    # the same host oracle with sane bounds walks a real trajectory; a short prefix keeps the Python loops cheap
    sane = [0.12, 0.20, (1.0 / 1.025) ** 2, 1.025 ** 2, 0.125 ** 2]
    x, niter, gamma = solve(kind, ref, lambda: LowpassHost(ref, 32, sane), 40.0, [0.0] * 32, sane[4], Options(150, 1e-14))
    assert niter == 150 and gamma <= sane[4]


# ------------------------------------------------------------------------------------------------ cutting_plane_tests.rs:328-379
class MyOracleFeas2:  # feasible when x + y <= 3; OracleFeas::update is the default no-op (src/cutting_plane.rs:125)
    def assess_feas(self, xc):
        fj = xc[0] + xc[1] - 3.0
        if fj > 0.0:
            return np.array([1.0, 1.0]), SingleCut(fj)
        return None

    def update(self, gamma):
        pass


@pytest.mark.parametrize("kind", SPACES)
def test_bsearch_adaptor_suite(ref, kind):
    def run(k):
        adaptor = BSearchAdaptor(MyOracleFeas2(), make_space(k, ref, 10.0, [0.0, 0.0]), Options())
        assert np.array_equal(adaptor.space.xc(), [0.0, 0.0])                                          # :370-379
        feasible, niter = bsearch(adaptor, (-100.0, 100.0), Options(2000, 1e-8))                       # :358-368
        return feasible, niter, adaptor.space.xc()
    feasible, niter, x = run(kind)
    assert feasible
    if kind == "gpu":
        f2, n2, x2 = run("checker")
        assert (feasible, niter) == (f2, n2) and np.array_equal(x, x2)


# ------------------------------------------------------------------------------------------------ benches/ellipsoid.rs
class BenchOracle:  # benches/ellipsoid.rs:10-24: gradient = 2*xc, f = xc.dot(xc)
    def assess_optim(self, xc, gamma):
        grad = 2.0 * xc
        f = 0.0
        for v in xc:
            f += v * v
        if f < gamma:
            return (grad, SingleCut(f)), True, f
        return (grad, SingleCut(f)), False, gamma


@pytest.mark.parametrize("kind", SPACES)
def test_reference_criterion_bench_configuration(ref, kind):
    """the reference's own benchmark (dimensions 10 / 50 / 100, Ell::new_with_scalar(10, 0), gamma = -inf, default Options)
    starts at the minimiser with a zero gradient: omega = 0, a 0/0 in the cut formulas, Success with NaN state, and the
    loop ends at iteration 0 on tsq < tol — the same on both spaces"""
    for dim in (10, 50, 100):                                                                          # :29-44
        x, niter, gamma = solve(kind, ref, BenchOracle, 10.0, [0.0] * dim, -INF, Options())
        assert x is None and niter == 0 and gamma == -INF


# ------------------------------------------------------------------------------------------------ examples/quickcheck_tests.rs
@pytest.mark.parametrize("kind", SPACES)
def test_quickcheck_properties(ref, kind):
    """the reference's property tests (examples/quickcheck_tests.rs:62-262) with its generators (n = 1..4, kappa = |v| % 10 +
    0.001), 100 seeded draws per property, on both spaces; on the GPU every draw must also reproduce the checker's status,
    kappa and tsq exactly"""
    rng = np.random.default_rng(20240)

    def draw():
        return int(rng.integers(1, 5)), float(abs(rng.normal(0.0, 50.0)) % 10.0 + 0.001)

    def step(n, kappa, update, grad, cut):
        spaces = [make_space(kind, ref, kappa, np.zeros(n))] + ([CheckerSpace(ref, kappa, np.zeros(n))] if kind == "gpu" else [])
        res = []
        for sp in spaces:
            st = getattr(sp, update)((np.full(n, grad), cut))
            kap = sp.kappa if hasattr(sp, "kappa") else sp.s.state()["kappa"]
            res.append((int(st), float(kap), float(sp.tsq())))
        if kind == "gpu":
            assert res[0][0] == res[1][0] and np.array_equal(res[0][1:], res[1][1:], equal_nan=True), res
        return res[0]

    for _ in range(100):
        n, kappa = draw()
        st, kap, _ = step(n, kappa, "update_central_cut", 0.01, SingleCut(0.0))              # :74-86
        assert st != CutStatus.Success or kap > 0.0
        n, kappa = draw()
        tsq = kappa * (n * 0.01 * 0.01)                                                        # M = I
        st, kap, _ = step(n, kappa, "update_bias_cut", 0.01, SingleCut(abs(math.sqrt(tsq) * 0.5)))   # :88-104
        assert st != CutStatus.Success or kap > 0.0
        n, kappa = draw()
        _, kap, _ = step(n, kappa, "update_central_cut", 0.1, SingleCut(0.0))                # :106-114
        assert kap > 0.0
        n, kappa = draw()
        st, kap, _ = step(n, kappa, "update_central_cut", 0.01, ParallelCut(0.0, 0.05))      # :163-175
        assert st != CutStatus.Success or kap > 0.0
        n, kappa = draw()
        g = np.full(n, 0.1)
        tsq = kappa * float(sum(v * v for v in g))
        st, _, _ = step(n, kappa, "update_bias_cut", 0.1, SingleCut(math.sqrt(tsq) * 1.5))   # :177-188
        assert st in (CutStatus.NoEffect, CutStatus.NoSoln)
    sp = make_space(kind, ref, 1.0, np.zeros(4))                                               # :190-201
    for _ in range(5):
        st = sp.update_central_cut((np.full(4, 0.1), SingleCut(0.0)))
        kap = sp.kappa if hasattr(sp, "kappa") else sp.s.state()["kappa"]
        assert st != CutStatus.Success or kap > 0.0
    for n in range(1, 9):                                                                       # :132-161 (EllCalc level)
        assert min(ref.calc(n, 1, tsq=0.1)[1]) > 0.0
        assert min(ref.calc(n, 0, 0.5, tsq=1.0)[1]) > 0.0
        assert min(ref.calc(n, 0, 0.3, 0.6, tsq=1.0)[1]) > 0.0


# ------------------------------------------------------------------------------------------------ examples/portfolio_optimization.rs
class PortfolioOracle:  # :24-74
    def __init__(self, expected_returns, risk_matrix, max_risk, budget):
        self.r, self.q, self.max_risk, self.budget = np.asarray(expected_returns, float), np.asarray(risk_matrix, float), max_risk, budget

    def assess_optim(self, xc, gamma):
        n = len(xc)
        grad = np.zeros(n)
        viol = 0.0
        for i in range(n):
            viol += xc[i]
        viol -= self.budget
        if viol > 0.0:
            grad[:] = 1.0
            return (grad, SingleCut(viol)), False, gamma
        risk = 0.0
        for i in range(n):
            for j in range(n):
                risk += xc[i] * self.q[i, j] * xc[j]
        rv = risk - self.max_risk
        if rv > 0.0:
            for i in range(n):
                s = 0.0
                for j in range(n):
                    s += (self.q[i, j] + self.q[j, i]) * xc[j]
                grad[i] = s
            return (grad, SingleCut(rv)), False, gamma
        ret = 0.0
        for a, b in zip(self.r, xc):
            ret += a * b
        obj = -ret
        if obj < gamma:
            return (-self.r, SingleCut(0.0)), True, obj
        return (grad, SingleCut(0.0)), False, gamma   # zero gradient: omega = 0, the cut formulas see 0/0 (as in the reference)


@pytest.mark.parametrize("kind", SPACES)
def test_portfolio_example(ref, kind):  # :76-95: 4 assets, identity risk, Options::new(1000, 1e-8); no asserts in the reference
    make = lambda: PortfolioOracle([0.08, 0.12, 0.10, 0.06], np.eye(4), 0.02, 1.0)
    x, niter, gamma = solve(kind, ref, make, 1.0, [0.0] * 4, INF, Options(1000, 1e-8))
    assert x is not None and niter <= 1000 and gamma <= 0.0
