"""GPU parity tests, large-n mode: ell.Ell (fused TMA sweep behind the C ABI) against the CPU checker, stepping the
same cuts.  reduce_mode 0 must be bit-identical (grad_t and omega are summed in the reference's order);
reduce_mode k>0 is held to north_star's 1e-9 relative."""
import ctypes as C
import math

import numpy as np
import pytest

from ellalgo_rs_b200 import _abi as A
from ellalgo_rs_b200 import workloads as W
from ellalgo_rs_b200.cutting_plane import (BSearchAdaptor, CutStatus, Options, ParallelCut, SingleCut, bsearch,
                                           cutting_plane_feas, cutting_plane_optim, cutting_plane_optim_q)
from ellalgo_rs_b200.ell import Ell
from parity import REL_TOL, rel_err

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------- reference unit tests, src/ell.rs:246-354
def test_ell_known_answers():
    g = 0.5 * np.ones(4)
    e = Ell.new_with_scalar(0.01, np.zeros(4))
    assert e.kappa == 0.01 and e.tsq() == 0.0 and np.array_equal(e.mq, np.eye(4)) and np.array_equal(e.xc(), np.zeros(4))
    assert e.update_central_cut((g, SingleCut(0.0))) == CutStatus.Success
    assert np.array_equal(e.xc(), -0.01 * np.ones(4))
    assert np.array_equal(e.mq, np.eye(4) - 0.1 * np.ones((4, 4)))
    assert e.kappa == pytest.approx(0.16 / 15, abs=1e-6) and e.tsq() == pytest.approx(0.01, abs=1e-6)

    e = Ell.new_with_scalar(0.01, np.zeros(4))
    assert e.update_bias_cut((g, SingleCut(0.05))) == CutStatus.Success
    assert e.xc()[0] == pytest.approx(-0.03, abs=1e-6) and e.mq[0, 0] == pytest.approx(0.8, abs=1e-6)
    assert e.kappa == pytest.approx(0.008, abs=1e-6)

    e = Ell.new_with_scalar(0.01, np.zeros(4))
    assert e.update_central_cut((g, ParallelCut(0.0, 0.05))) == CutStatus.Success
    assert np.array_equal(e.mq, np.eye(4) - 0.2 * np.ones((4, 4))) and e.kappa == pytest.approx(0.012, abs=1e-6)

    e = Ell.new_with_scalar(0.01, np.zeros(4))
    assert e.update_bias_cut((g, ParallelCut(0.01, 0.04))) == CutStatus.Success
    assert e.xc()[0] == pytest.approx(-0.0116, abs=1e-6) and e.mq[0, 0] == pytest.approx(1 - 0.232, abs=1e-6)
    assert e.kappa == pytest.approx(0.01232, abs=1e-6)

    e = Ell.new_with_scalar(0.01, np.zeros(4))  # :294-314
    assert e.update_bias_cut((g, ParallelCut(-0.04, 0.0625))) == CutStatus.Success
    assert np.allclose(e.mq, np.eye(4), atol=1e-6) and e.kappa == pytest.approx(0.01, abs=1e-6)
    e = Ell.new_with_scalar(0.01, np.zeros(4))
    assert e.update_q((g, ParallelCut(-0.04, 0.0625))) == CutStatus.NoEffect
    assert np.array_equal(e.mq, np.eye(4)) and e.kappa == 0.01 and np.array_equal(e.xc(), np.zeros(4))

    e = Ell.new_with_scalar(0.01, np.zeros(4), no_defer_trick=True)  # :341-354
    assert e.update_central_cut((g, SingleCut(0.0))) == CutStatus.Success
    assert e.kappa == 1.0


def test_ell_construct_and_accessors():  # src/ell_test.rs:12-20, 131-155 on the GPU search space
    e = Ell.new_with_scalar(0.01, np.zeros(4))
    assert e.kappa == 0.01 and e.tsq() == 0.0 and np.array_equal(e.mq, np.eye(4)) and np.array_equal(e.xc(), np.zeros(4))
    e = Ell(1.0, np.zeros(3), diag=np.array([1.0, 2.0, 3.0]))            # Ell::new(val, xc)
    assert e.kappa == 1.0 and np.array_equal(e.mq, np.diag([1.0, 2.0, 3.0])) and e.tsq() == 0.0
    e = Ell.new_with_scalar(1.0, np.zeros(2))
    e.set_xc(np.array([1.0, 2.0]))
    assert np.array_equal(e.xc(), [1.0, 2.0]) and e.tsq() == 0.0


# ---------------------------------------------------------------- stepwise trajectories against the checker
def _cuts(n, steps, seed):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(steps):
        g = rng.uniform(-1, 1, n) / math.sqrt(n)
        kind = k % 6
        if kind == 0: out.append((A.UPDATE_BIAS, g, ParallelCut(0.01, 0.10)))
        elif kind == 1: out.append((A.UPDATE_CENTRAL, g, SingleCut(0.0)))
        elif kind == 2: out.append((A.UPDATE_BIAS, g, SingleCut(1e6)))          # NoSoln: state must not move
        elif kind == 3: out.append((A.UPDATE_BIAS, g, SingleCut(0.02)))
        elif kind == 4: out.append((A.UPDATE_Q, g, ParallelCut(-0.5, None)))
        else: out.append((A.UPDATE_CENTRAL, g, ParallelCut(0.0, 0.05)))
    return out


def _ref_update(space, which, g, cut):
    if isinstance(cut, SingleCut):
        return space.update(which, g, cut.beta)
    return space.update(which, g, cut.beta0, cut.beta1, single=False)


def _gpu_update(e, which, g, cut):
    return [e.update_bias_cut, e.update_central_cut, e.update_q][which]((g, cut))


def _start(n, start, seed):
    rng = np.random.default_rng(seed + 1000)
    xc = rng.uniform(-1, 1, n)
    if start == "scalar":
        return dict(kappa=2.5, xc=xc)
    if start == "diag":
        return dict(kappa=1.0, xc=xc, diag=rng.uniform(0.5, 2.0, n))
    a = rng.uniform(-1, 1, (n, n)) / n
    return dict(kappa=1.0, xc=xc, mq=np.eye(n) + a + a.T)


@pytest.mark.parametrize("n,start", [(2, "scalar"), (5, "matrix"), (33, "diag"), (64, "scalar"), (100, "matrix"),
                                     (257, "matrix"), (1000, "scalar")])
@pytest.mark.parametrize("no_defer", [False, True])
def test_trajectory_bit_identical(ref, n, start, no_defer):
    kw = _start(n, start, n)
    e = Ell(kw["kappa"], kw["xc"], mq=kw.get("mq"), diag=kw.get("diag"), no_defer_trick=no_defer)
    s = ref.Space(A.KIND_ELL, n, kappa=kw["kappa"], mq=kw.get("mq"), diag=kw.get("diag"), xc=kw["xc"], no_defer_trick=no_defer)
    for k, (which, g, cut) in enumerate(_cuts(n, 9, n)):
        st_g, st_r = _gpu_update(e, which, g, cut), _ref_update(s, which, g, cut)
        r = s.state()
        assert int(st_g) == st_r, f"step {k}: status"
        assert e.tsq() == r["tsq"] and e.kappa == r["kappa"], f"step {k}: tsq/kappa not bit-identical"
        assert np.array_equal(e.xc(), r["xc"]), f"step {k}: xc"
        if k in (3, 8):
            assert np.array_equal(e.mq, r["mq"]), f"step {k}: mq"


@pytest.mark.parametrize("mode", [2, 4, 7])
def test_trajectory_blocked_reduction(ref, mode):
    n = 300
    kw = _start(n, "matrix", 5)
    e = Ell(kw["kappa"], kw["xc"], mq=kw["mq"], reduce_mode=mode)
    s = ref.Space(A.KIND_ELL, n, kappa=kw["kappa"], mq=kw["mq"], xc=kw["xc"])
    for which, g, cut in _cuts(n, 8, 77):
        assert int(_gpu_update(e, which, g, cut)) == _ref_update(s, which, g, cut)
    r = s.state()
    assert rel_err(e.tsq(), r["tsq"]).max() <= REL_TOL and rel_err(e.kappa, r["kappa"]).max() <= REL_TOL
    assert rel_err(e.xc(), r["xc"]).max() <= REL_TOL
    assert np.abs(e.mq - r["mq"]).max() <= REL_TOL * np.abs(r["mq"]).max()


def test_use_parallel_cut_flag_and_clone(ref):
    n = 48
    kw = _start(n, "scalar", 9)
    e = Ell(kw["kappa"], kw["xc"], use_parallel_cut=False)
    s = ref.Space(A.KIND_ELL, n, kappa=kw["kappa"], xc=kw["xc"], use_parallel_cut=False)
    cuts = _cuts(n, 4, 3)
    for which, g, cut in cuts[:2]:
        assert int(_gpu_update(e, which, g, cut)) == _ref_update(s, which, g, cut)
    c = e.clone()                      # #[derive(Clone)]: an independent copy, pending rank-1 included
    which, g, cut = cuts[3]
    _gpu_update(c, which, g, cut)
    r = s.state()
    assert np.array_equal(e.xc(), r["xc"]) and e.kappa == r["kappa"] and np.array_equal(e.mq, r["mq"])
    _ref_update(s, which, g, cut)
    r = s.state()
    assert np.array_equal(c.xc(), r["xc"]) and c.kappa == r["kappa"] and np.array_equal(c.mq, r["mq"])
    e.set_xc(np.arange(n, dtype=float))
    assert np.array_equal(e.xc(), np.arange(n, dtype=float))


def test_bad_arguments():
    from ellalgo_rs_b200._lib import EllCudaError
    e = Ell(1.0, np.zeros(3), mq=np.arange(9.0).reshape(3, 3))      # non-symmetric: accepted, as Ell::new_with_matrix does
    assert np.array_equal(e.mq, np.arange(9.0).reshape(3, 3))
    with pytest.raises(EllCudaError):
        Ell(1.0, np.zeros(3), device=99)
    with pytest.raises(ValueError):
        Ell(1.0, np.zeros(3), mq=np.eye(4))
    e = Ell.new_with_scalar(1.0, np.zeros(3))
    with pytest.raises(ValueError):
        e.update_bias_cut((np.zeros(4), SingleCut(0.0)))


# ---------------------------------------------------------------- sharded slabs on one GPU (ranks emulated in-process)
@pytest.mark.parametrize("n,parts,mode", [(96, 2, 0), (130, 3, 0), (257, 4, 2), (1024, 8, 0)])
def test_row_sharded_slabs_peer_exchange(ref, n, parts, mode):
    """the NVLink peer-exchange path (ellcuda_exchange_attach_ptrs) with the ranks emulated on one GPU: all handles on
    ONE stream, begin for every rank before any finish, so no kernel ever waits for a later launch"""
    import torch
    kw = _start(n, "matrix", 31)
    s = ref.Space(A.KIND_ELL, n, kappa=kw["kappa"], mq=kw["mq"], xc=kw["xc"])
    bounds = [n * p // parts for p in range(parts + 1)]
    stream = torch.cuda.Stream()
    shards = [Ell(kw["kappa"], kw["xc"], mq=kw["mq"], rows=(bounds[p], bounds[p + 1]), reduce_mode=mode, stream=stream.cuda_stream)
              for p in range(parts)]
    blocks = [sh.exchange_block_ptr() for sh in shards]
    for p, sh in enumerate(shards):
        sh.exchange_attach_ptrs(blocks, p)
    with torch.cuda.stream(stream):
        for which, g, cut in _cuts(n, 8, 15):
            gd = torch.from_numpy(g).cuda()
            for sh in shards:
                sh.update_begin(which, gd.data_ptr(), cut)
            for sh in shards:
                sh.update_finish()
            st_r = _ref_update(s, which, g, cut)
            r = s.state()
            for sh in shards:
                assert int(sh.last_status()) == st_r
                if mode == 0:
                    assert sh.tsq() == r["tsq"] and sh.kappa == r["kappa"] and np.array_equal(sh.xc(), r["xc"])
                else:
                    assert rel_err(sh.tsq(), r["tsq"]).max() <= REL_TOL and rel_err(sh.xc(), r["xc"]).max() <= REL_TOL
    full = np.concatenate([sh.mq for sh in shards], axis=0)
    r = s.state()
    assert np.array_equal(full, r["mq"]) if mode == 0 else np.abs(full - r["mq"]).max() <= REL_TOL * np.abs(r["mq"]).max()


@pytest.mark.parametrize("n,parts,mode", [(96, 2, 0), (130, 3, 0), (257, 4, 2)])
def test_row_sharded_slabs(ref, n, parts, mode):
    import torch
    kw = _start(n, "matrix", 21)
    s = ref.Space(A.KIND_ELL, n, kappa=kw["kappa"], mq=kw["mq"], xc=kw["xc"])
    bounds = [n * p // parts for p in range(parts + 1)]
    shards = [Ell(kw["kappa"], kw["xc"], mq=kw["mq"], rows=(bounds[p], bounds[p + 1]), reduce_mode=mode) for p in range(parts)]
    views = [torch.as_tensor(sh.grad_t_view(), device="cuda") for sh in shards]
    for which, g, cut in _cuts(n, 7, 5):
        gd = torch.from_numpy(g).cuda()
        for sh in shards:
            sh.update_begin(which, gd.data_ptr(), cut)
        for sh in shards:
            sh.sync()
        for p in range(parts):          # the allgather of grad_t slabs
            for q in range(parts):
                if p != q:
                    views[q][bounds[p]:bounds[p + 1]].copy_(views[p][bounds[p]:bounds[p + 1]])
        torch.cuda.synchronize()
        for sh in shards:
            sh.update_finish()
        st_r = _ref_update(s, which, g, cut)
        r = s.state()
        for sh in shards:
            assert int(sh.last_status()) == st_r
            if mode == 0:
                assert sh.tsq() == r["tsq"] and sh.kappa == r["kappa"] and np.array_equal(sh.xc(), r["xc"])
            else:
                assert rel_err(sh.tsq(), r["tsq"]).max() <= REL_TOL and rel_err(sh.xc(), r["xc"]).max() <= REL_TOL
    full = np.concatenate([sh.mq for sh in shards], axis=0)
    r = s.state()
    if mode == 0:
        assert np.array_equal(full, r["mq"])
    else:
        assert np.abs(full - r["mq"]).max() <= REL_TOL * np.abs(r["mq"]).max()


# ---------------------------------------------------------------- the reference's drivers over the GPU search space
class _Example1:  # src/example1.rs:8-30
    def assess_optim(self, xc, gamma):
        x, y = xc
        f0 = x + y
        f1 = f0 - 3.0
        if f1 > 0.0: return (np.array([1.0, 1.0]), SingleCut(f1)), False, gamma
        f2 = -x + y + 1.0
        if f2 > 0.0: return (np.array([-1.0, 1.0]), SingleCut(f2)), False, gamma
        f3 = gamma - f0
        if f3 > 0.0: return (np.array([-1.0, -1.0]), SingleCut(f3)), False, gamma
        return (np.array([-1.0, -1.0]), SingleCut(0.0)), True, f0


class _Example3:  # src/example3.rs:10-60
    def __init__(self): self.idx, self.target = -1, -1e100
    def assess_feas(self, xc):
        x, y = xc
        for _ in range(4):
            self.idx = (self.idx + 1) % 4
            fj = [-x - 1.0, -y - 2.0, x + y - 1.0, 2.0 * x - 3.0 * y - self.target][self.idx]
            if fj > 0.0:
                return np.array([[-1.0, 0.0], [0.0, -1.0], [1.0, 1.0], [2.0, -3.0]][self.idx]), SingleCut(fj)
        return None
    def update(self, gamma): self.target = gamma


def test_driver_example1_on_gpu_space():  # src/example1.rs:39-70
    e = Ell.new_with_scalar(10.0, np.zeros(2))
    xbest, niter, gamma = cutting_plane_optim(_Example1(), e, -math.inf, Options(tolerance=1e-10))
    assert xbest is not None and niter == 25
    e = Ell.new(np.array([10.0, 10.0]), np.array([100.0, 100.0]))
    assert cutting_plane_optim(_Example1(), e, -math.inf, Options())[0] is None
    e = Ell.new(np.array([10.0, 10.0]), np.zeros(2))
    assert cutting_plane_optim(_Example1(), e, 100.0, Options())[0] is None


def test_driver_bsearch_on_gpu_space():  # src/example3.rs:68-85 — needs clone() and set_xc()
    e = Ell.new_with_scalar(100.0, np.zeros(2))
    adaptor = BSearchAdaptor(_Example3(), e, Options(tolerance=1e-8))
    feasible, niter = bsearch(adaptor, (-100.0, 100.0), Options(tolerance=1e-8))
    assert (feasible, niter) == (True, 34)


def test_driver_feas_and_q_fixtures():  # tests/cutting_plane_tests.rs:128-146, 266-327
    class AlwaysCut:
        def assess_feas(self, xc): return np.array([1.0, 1.0]), SingleCut(1.0)
        def assess_optim_q(self, xc, gamma, retry): return (np.array([1.0, 1.0]), SingleCut(1.0)), False, xc.copy(), True, gamma
    assert cutting_plane_feas(AlwaysCut(), Ell.new_with_scalar(10.0, np.zeros(2)), Options(200, 1e-20)) == (None, 2)
    x, niter, _ = cutting_plane_optim_q(AlwaysCut(), Ell.new_with_scalar(10.0, np.zeros(2)), 0.0, Options(5, 1e-20))
    assert x is None and niter == 2


# ---------------------------------------------------------------- the benchmark shape itself
@pytest.mark.parametrize("n,mode", [(4096, 0), (16384, 0), (16384, 8)])
def test_large_n_against_checker(ref, n, mode):
    """C5: M = I, ParallelCut(0.01, Some(0.10)) via update_bias_cut, cuts from the seeded generator."""
    steps = 3 if n <= 4096 else 2
    e = Ell.new_with_scalar(1.0, np.zeros(n), reduce_mode=mode)
    mq, xc = np.eye(n), np.zeros(n)
    kappa, tsq = C.c_double(1.0), C.c_double(0.0)
    for k in range(steps):
        g = W.large_n_grad(n, k)
        st = e.update_bias_cut((g, ParallelCut(*W.LARGE_N_CUT)))
        st_r = ref.lib().ellref_ell_update_inplace(n, mq.ctypes.data, xc.ctypes.data, C.byref(kappa), C.byref(tsq), 0, 1, 0,
                                                   g.ctypes.data, W.LARGE_N_CUT[0], 1, W.LARGE_N_CUT[1], None)
        assert int(st) == st_r == A.SUCCESS
        if mode == 0:
            assert e.tsq() == tsq.value and e.kappa == kappa.value and np.array_equal(e.xc(), xc)
        else:
            assert rel_err(e.tsq(), tsq.value).max() <= REL_TOL and rel_err(e.kappa, kappa.value).max() <= REL_TOL
            assert np.abs(e.xc() - xc).max() <= REL_TOL * np.abs(xc).max()
    got = e.mq
    if mode == 0:
        assert np.array_equal(got, mq)
    else:
        assert np.abs(got - mq).max() <= REL_TOL
    assert np.array_equal(got, got.T)  # stays bitwise symmetric, like the reference's mirrored writes


# ---------------------------------------------------------------- EllStable as a single host-driven search space
def test_ellstable_known_answers():  # src/ell_stable.rs:225-298
    from ellalgo_rs_b200.ell import EllStable
    g = 0.5 * np.ones(4)
    e = EllStable.new_with_scalar(0.01, np.zeros(4))
    assert e.kappa == 0.01 and e.tsq() == 0.0 and np.array_equal(e.xc(), np.zeros(4))
    assert e.update_central_cut((g, SingleCut(0.0))) == CutStatus.Success
    assert np.array_equal(e.xc(), -0.01 * np.ones(4)) and e.kappa == pytest.approx(0.16 / 15, abs=1e-6)
    assert e.tsq() == pytest.approx(0.01, abs=1e-6)
    e = EllStable.new_with_scalar(0.01, np.zeros(4))
    assert e.update_bias_cut((g, SingleCut(0.05))) == CutStatus.Success
    assert e.xc()[0] == pytest.approx(-0.03, abs=1e-6) and e.kappa == pytest.approx(0.008, abs=1e-6)
    e = EllStable.new_with_scalar(0.01, np.zeros(4))
    assert e.update_central_cut((g, ParallelCut(0.0, 0.05))) == CutStatus.Success
    assert np.array_equal(e.xc(), -0.01 * np.ones(4)) and e.kappa == pytest.approx(0.012, abs=1e-6)
    e = EllStable.new_with_scalar(0.01, np.zeros(4))
    assert e.update_bias_cut((g, ParallelCut(0.01, 0.04))) == CutStatus.Success
    assert e.xc()[0] == pytest.approx(-0.0116, abs=1e-6) and e.kappa == pytest.approx(0.01232, abs=1e-6)
    e = EllStable.new_with_scalar(0.01, np.zeros(4))
    assert e.update_q((g, ParallelCut(-0.04, 0.0625))) == CutStatus.NoEffect
    assert np.array_equal(e.xc(), np.zeros(4)) and e.kappa == 0.01


@pytest.mark.parametrize("n,start", [(2, "scalar"), (7, "diag"), (33, "upper"), (64, "upper"), (64, "scalar"),
                                     (65, "upper"), (100, "diag"), (257, "upper"), (700, "scalar")])   # > 64: one-CTA kernel
def test_ellstable_trajectory_bit_identical(ref, n, start):
    from ellalgo_rs_b200.ell import EllStable
    rng = np.random.default_rng(n)
    xc = rng.uniform(-1, 1, n)
    kw = dict(kappa=3.0, xc=xc)
    if start == "diag":
        kw.update(kappa=1.0, diag=rng.uniform(0.5, 2.0, n))
    elif start == "upper":
        kw.update(mq=np.eye(n) + np.triu(rng.uniform(-1e-3, 1e-3, (n, n)), 1))
    e = EllStable(kw["kappa"], xc, mq=kw.get("mq"), diag=kw.get("diag"))
    s = ref.Space(A.KIND_ELLSTABLE, n, kappa=kw["kappa"], mq=kw.get("mq"), diag=kw.get("diag"), xc=xc)
    for k, (which, g, cut) in enumerate(_cuts(n, 9, n + 5)):
        st_g, st_r = _gpu_update(e, which, g, cut), _ref_update(s, which, g, cut)
        r = s.state()
        assert int(st_g) == st_r, f"step {k}: status"
        assert e.tsq() == r["tsq"] and e.kappa == r["kappa"] and np.array_equal(e.xc(), r["xc"]), f"step {k}"
    got, want = e.mq, s.state()["mq"]
    assert np.array_equal(np.triu(got), np.triu(want))      # the factor; the strict lower triangle is scratch in the reference
    c = e.clone()
    which, g, cut = _cuts(n, 1, 99)[0]
    _gpu_update(c, which, g, cut)
    assert np.array_equal(e.xc(), s.state()["xc"])          # the clone is independent


def test_ellstable_driver_quasicvx():  # src/quasicvx.rs:95-105 test_feasible_stable
    from ellalgo_rs_b200.ell import EllStable

    class Quasicvx:
        def __init__(self): self.idx = -1
        def assess_optim(self, xc, gamma):
            sqrtx, logy = xc
            for _ in range(2):
                self.idx = (self.idx + 1) % 2
                fv = sqrtx * sqrtx - logy if self.idx == 0 else -sqrtx + gamma * math.exp(logy)
                if fv > 0.0:
                    g = np.array([2.0 * sqrtx, -1.0]) if self.idx == 0 else np.array([-1.0, gamma * math.exp(logy)])
                    return (g, SingleCut(fv)), False, gamma
            return (np.array([-1.0, sqrtx]), SingleCut(0.0)), True, sqrtx / math.exp(logy)
    x, niter, _ = cutting_plane_optim(Quasicvx(), EllStable.new(np.array([10.0, 10.0]), np.zeros(2)), 0.0, Options(2000, 1e-8))
    assert x is not None and niter == 49
    with pytest.raises(Exception):
        EllStable.new_with_scalar(1.0, np.zeros(2049))     # one-CTA kernel: n <= 2048


# ---------------------------------------------------------------- §8f rank 4: the streaming GEMV on its own
def test_dot_mv_and_its_users_on_the_device(ref):
    """Arr::dot_mv through ellcuda_dot_mv, then power_iteration (src/power_iteration.rs:178-229 → 22/21/23/23) and
    conjugate_gradient (src/conjugate_gradient.rs) driven by it: bit-identical to the checker."""
    from ellalgo_rs_b200 import linalg as LA
    from test_oracle_pins import POWER_A
    a = Ell.from_covariance(POWER_A, np.zeros(3))
    x0 = np.array([0.3, 0.5, 0.4])
    for v, f, tol, pinned in ((1, LA.power_iteration, 1e-7, 22), (4, LA.power_iteration4, 1e-7, 21),
                              (2, LA.power_iteration2, 1e-14, 23), (3, LA.power_iteration3, 1e-14, 23)):
        e, it, x = f(a, x0, LA.Options(2000, tol))
        er, itr, xr = ref.power_iteration(v, POWER_A, x0, 2000, tol)
        assert it == itr == pinned and e == er and np.array_equal(x, xr)
    cg = Ell.from_covariance(np.array([[4.0, 1.0], [1.0, 3.0]]), np.zeros(2))
    x = LA.conjugate_gradient(cg, [1.0, 2.0], None, 1e-5, 1000)
    assert np.array_equal(x, ref.conjugate_gradient(np.array([[4.0, 1.0], [1.0, 3.0]]), [1.0, 2.0], None, 1e-5, 1000)[0])

    n = 777                                     # a size that is no multiple of the tile, dense SPD
    rng = np.random.default_rng(4)
    b = rng.uniform(-1, 1, (n, n)) / math.sqrt(n)
    spd = b @ b.T + np.eye(n)
    spd = (spd + spd.T) / 2
    m = Ell.from_covariance(spd, np.zeros(n))
    v = rng.uniform(-1, 1, n)
    want = np.array([float(np.cumsum(r * v)[-1]) for r in spd])
    assert np.array_equal(m.dot_mv(v), want)    # left-to-right row sums, bit for bit
    rhs = rng.uniform(-1, 1, n)
    x = LA.conjugate_gradient(m, rhs, None, 1e-9, 500)
    xr, it = ref.conjugate_gradient(spd, rhs, None, 1e-9, 500)
    assert it >= 0 and np.array_equal(x, xr)
    x0 = rng.uniform(0, 1, n)
    e, it, xv = LA.power_iteration2(m, x0, LA.Options(300, 1e-10))
    er, itr, xvr = ref.power_iteration(2, spd, x0, 300, 1e-10)
    assert (it, e) == (itr, er) and np.array_equal(xv, xvr)
    # the same iterations with the vectors resident on the GPU (ellcuda_seq_sum for every ordered sum): identical bits
    x = LA.conjugate_gradient(m, rhs, None, 1e-9, 500, device=True)
    assert np.array_equal(x, xr)
    e, it, xv = LA.power_iteration2(m, x0, LA.Options(300, 1e-10), device=True)
    assert (it, e) == (itr, er) and np.array_equal(xv, xvr)
    for variant, f, tol in ((1, LA.power_iteration, 1e-9), (4, LA.power_iteration4, 1e-9), (3, LA.power_iteration3, 1e-12)):
        e, it, xv = f(m, x0, LA.Options(200, tol), device=True)
        er2, itr2, xvr2 = ref.power_iteration(variant, spd, x0, 200, tol)
        assert (it, e) == (itr2, er2) and np.array_equal(xv, xvr2), variant
    import torch
    t = torch.from_numpy(rng.uniform(-1, 1, 5000)).cuda()
    assert m.seq_sum(t) == float(np.cumsum(t.cpu().numpy())[-1])
    # dot_mv with a deferred rank-1 pending: the product must be with the updated matrix
    g = W.large_n_grad(n, 0)
    s = ref.Space(A.KIND_ELL, n, kappa=1.0, mq=spd)
    m.update_bias_cut((g, ParallelCut(0.001, 0.01)))
    s.update(0, g, 0.001, 0.01, single=False)
    want = np.array([float(np.cumsum(r * v)[-1]) for r in s.state()["mq"]])
    assert np.array_equal(m.dot_mv(v), want)
