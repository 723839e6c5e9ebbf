"""GPU tests of the boundary pieces the reference exposes beyond the plain constructors:

  * `ellip.no_defer_trick = true` / `helper.use_parallel_cut` written AFTER construction   src/ell.rs:10, 342-354; ell_calc.rs:630
  * `Ell::new_with_matrix` with a NON-symmetric matrix                                   src/ell.rs:31-41, 102, 117-128
  * one `Ell` row-sharded over several devices of ONE process (`ndev, devs`)  ellcuda_create_sharded
  * one batch spread over several devices of one process                                 ellcuda_batch_solve_multi

All against the CPU checker, bit for bit unless a reduce_mode > 0 says otherwise."""
import ctypes as C
import math

import numpy as np
import pytest

from ellalgo_rs_b200 import _abi as A
from ellalgo_rs_b200 import batch
from ellalgo_rs_b200 import workloads as W
from ellalgo_rs_b200._lib import EllCudaError, lib
from ellalgo_rs_b200.cutting_plane import CutStatus, ParallelCut, SingleCut
from ellalgo_rs_b200.ell import Ell, EllStable
from ellalgo_rs_b200.problem import BatchProblem
from parity import REL_TOL, assert_parity, rel_err
from test_stream_gpu import _cuts, _gpu_update, _ref_update, _start

pytestmark = pytest.mark.gpu


def _ndev():
    return lib().ellcuda_device_count()


# ---------------------------------------------------------------- knobs written after construction
def test_no_defer_trick_set_after_new():  # src/ell.rs:341-354, as the reference writes it
    e = Ell.new_with_scalar(0.01, np.zeros(4))
    e.no_defer_trick = True
    assert e.update_central_cut((0.5 * np.ones(4), SingleCut(0.0))) == CutStatus.Success
    assert e.kappa == 1.0
    expected = (np.eye(4) - 0.1 * np.ones((4, 4))) * (0.16 / 15.0)
    assert np.allclose(e.mq, expected, rtol=0, atol=1e-6)


def test_knobs_toggled_mid_trajectory(ref):
    n = 70
    kw = _start(n, "matrix", 4)
    e = Ell(kw["kappa"], kw["xc"], mq=kw["mq"])
    s = ref.Space(A.KIND_ELL, n, kappa=kw["kappa"], mq=kw["mq"], xc=kw["xc"])
    for k, (which, g, cut) in enumerate(_cuts(n, 12, 8)):
        if k == 3:
            e.no_defer_trick = True
            s2 = ref.Space(A.KIND_ELL, n, kappa=s.state()["kappa"], mq=s.state()["mq"], xc=s.state()["xc"], no_defer_trick=True)
            s = s2
        if k == 6:
            e.use_parallel_cut = False
            r = s.state()
            s = ref.Space(A.KIND_ELL, n, kappa=r["kappa"], mq=r["mq"], xc=r["xc"], no_defer_trick=True, use_parallel_cut=False)
        if k == 9:
            e.no_defer_trick = False
            r = s.state()
            s = ref.Space(A.KIND_ELL, n, kappa=r["kappa"], mq=r["mq"], xc=r["xc"], no_defer_trick=False, use_parallel_cut=False)
        assert int(_gpu_update(e, which, g, cut)) == _ref_update(s, which, g, cut), f"step {k}"
        r = s.state()
        assert e.tsq() == r["tsq"] and e.kappa == r["kappa"] and np.array_equal(e.xc(), r["xc"]), f"step {k}"
    assert np.array_equal(e.mq, s.state()["mq"])


def test_ellstable_has_no_defer_switch():
    e = EllStable.new_with_scalar(1.0, np.zeros(3))
    with pytest.raises((EllCudaError, ValueError)):
        e.no_defer_trick = True
    e.use_parallel_cut = False            # the helper's switch exists on both spaces (src/ell_stable.rs:14)


# ---------------------------------------------------------------- non-symmetric new_with_matrix
def _asym(n, seed):
    rng = np.random.default_rng(seed)
    return np.eye(n) + rng.uniform(-1, 1, (n, n)) / (2 * n)


@pytest.mark.parametrize("n", [3, 33, 65, 300])
@pytest.mark.parametrize("no_defer", [False, True])
def test_asymmetric_matrix_single_handle(ref, n, no_defer):
    mq = _asym(n, n)
    xc = np.random.default_rng(n + 1).uniform(-1, 1, n)
    e = Ell.new_with_matrix(1.5, mq, xc, no_defer_trick=no_defer)
    s = ref.Space(A.KIND_ELL, n, kappa=1.5, mq=mq, xc=xc, no_defer_trick=no_defer)
    assert np.array_equal(e.mq, mq)                       # pub field: the matrix as given
    cuts = _cuts(n, 9, n + 2)
    # lead with two failing cuts: the rows as given are multiplied (src/ell.rs:102) and nothing is mirrored yet
    cuts = [(A.UPDATE_BIAS, cuts[0][1], SingleCut(1e6)), (A.UPDATE_BIAS, cuts[1][1], ParallelCut(5.0, 4.0))] + cuts
    for k, (which, g, cut) in enumerate(cuts):
        st_g, st_r = _gpu_update(e, which, g, cut), _ref_update(s, which, g, cut)
        r = s.state()
        assert int(st_g) == st_r, f"step {k}"
        assert e.tsq() == r["tsq"] and e.kappa == r["kappa"] and np.array_equal(e.xc(), r["xc"]), f"step {k}"
        if k in (1, 2, 5, len(cuts) - 1):
            assert np.array_equal(e.mq, r["mq"]), f"step {k}: mq"
        if k == 1:
            c = e.clone()                                 # a clone taken before the first success carries the raw matrix too
            sc = s.clone()
    which, g, cut = cuts[2]
    assert int(_gpu_update(c, which, g, cut)) == _ref_update(sc, which, g, cut)
    assert np.array_equal(c.mq, sc.state()["mq"]) and np.array_equal(c.xc(), sc.state()["xc"])


def test_asymmetric_matrix_dot_mv(ref):
    n = 130
    mq = _asym(n, 5)
    e = Ell.new_with_matrix(1.0, mq, np.zeros(n))
    x = np.random.default_rng(3).uniform(-1, 1, n)
    y = np.zeros(n)
    ref.lib().ellref_dot_mv(mq.ctypes.data, x.ctypes.data, n, n, y.ctypes.data)
    assert np.array_equal(e.dot_mv(x), y)                 # rows as given, left-to-right sums


@pytest.mark.parametrize("n,oracle", [(2, "small"), (7, "bowl"), (32, "bowl"), (32, "lowpass"), (20, "lowpass")])
def test_asymmetric_matrix_batched(ref, n, oracle):
    cnt = 37
    rng = np.random.default_rng(n * 7 + len(oracle))
    mq0 = np.stack([_asym(n, 100 + i) * (40.0 if oracle == "lowpass" else 1.0) for i in range(cnt)])
    if oracle == "small":
        prob = W.profit(cnt)
        prob.diag0 = None
        mq0 = np.stack([np.array([[100.0, 3.0 * rng.uniform()], [-2.0 * rng.uniform(), 100.0]]) for _ in range(cnt)])
    elif oracle == "bowl":
        prob = W.bowl(cnt, n=n, max_iters=300)
    else:
        prob = W.lowpass(cnt, batch.lowpass_table(n), n=n, max_iters=400)
        prob.kappa0 = np.ones(1)
    prob.mq0 = np.ascontiguousarray(mq0)
    g, r = batch.solve_batch(prob), ref.solve(prob)
    assert_parity(g, r, exact=True, what=f"asymmetric mq0, {oracle} n={n}")
    assert g.num_iters.max() > 3


def test_asymmetric_matrix_sharded_slabs(ref):
    """row-sharded handles built from a non-symmetric matrix (ranks emulated on one GPU, one stream)"""
    import torch
    n, parts = 200, 4
    mq = _asym(n, 9)
    xc = np.zeros(n)
    s = ref.Space(A.KIND_ELL, n, kappa=1.0, mq=mq, xc=xc)
    bounds = [n * p // parts for p in range(parts + 1)]
    stream = torch.cuda.Stream()
    shards = [Ell(1.0, xc, mq=mq, rows=(bounds[p], bounds[p + 1]), stream=stream.cuda_stream) for p in range(parts)]
    blocks = [sh.exchange_block_ptr() for sh in shards]
    for p, sh in enumerate(shards):
        sh.exchange_attach_ptrs(blocks, p)
    assert np.array_equal(np.concatenate([sh.mq for sh in shards]), mq)
    with torch.cuda.stream(stream):
        cuts = _cuts(n, 6, 2)
        cuts = [(A.UPDATE_BIAS, cuts[0][1], SingleCut(1e6))] + cuts
        for which, g, cut in cuts:
            gd = torch.from_numpy(g).cuda()
            for sh in shards:
                sh.update_begin(which, gd.data_ptr(), cut)
            for sh in shards:
                sh.update_finish()
            st_r = _ref_update(s, which, g, cut)
            r = s.state()
            for sh in shards:
                assert int(sh.last_status()) == st_r and sh.tsq() == r["tsq"] and np.array_equal(sh.xc(), r["xc"])
    assert np.array_equal(np.concatenate([sh.mq for sh in shards]), s.state()["mq"])


# ---------------------------------------------------------------- one process, several devices (or several slabs on one)
@pytest.mark.parametrize("n", [96, 257, 1024])
@pytest.mark.parametrize("case", range(5))
def test_group_handle(ref, n, case):
    nd = _ndev()
    devices, mode = [([0], 0), ([0, 0], 0), ([0, 0, 0], 3), ([0, 1], 0), (list(range(min(nd, 8))), min(nd, 8))][case]
    if max(devices) >= nd or (case == 4 and nd < 3):
        pytest.skip("needs more GPUs in this process")
    kw = _start(n, "matrix", 17)
    e = Ell(kw["kappa"], kw["xc"], mq=kw["mq"], devices=devices, reduce_mode=mode)
    assert e.group_size == len(devices)
    s = ref.Space(A.KIND_ELL, n, kappa=kw["kappa"], mq=kw["mq"], xc=kw["xc"])
    exact = mode == 0
    for k, (which, g, cut) in enumerate(_cuts(n, 9, 23)):
        if k == 4:
            e.no_defer_trick = True
            r = s.state()
            s = ref.Space(A.KIND_ELL, n, kappa=r["kappa"], mq=r["mq"], xc=r["xc"], no_defer_trick=True)
        st_g, st_r = _gpu_update(e, which, g, cut), _ref_update(s, which, g, cut)
        r = s.state()
        assert int(st_g) == st_r, f"step {k}"
        if exact:
            assert e.tsq() == r["tsq"] and e.kappa == r["kappa"] and np.array_equal(e.xc(), r["xc"]), f"step {k}"
        else:
            assert rel_err(e.tsq(), r["tsq"]).max() <= REL_TOL and rel_err(e.xc(), r["xc"]).max() <= REL_TOL, f"step {k}"
        if k == 5:
            c = e.clone()
            sc = s.clone()
    full, r = e.mq, s.state()
    assert full.shape == (n, n)
    assert np.array_equal(full, r["mq"]) if exact else np.abs(full - r["mq"]).max() <= REL_TOL * np.abs(r["mq"]).max()
    # the clone is a group of its own, attached to itself
    which, g, cut = _cuts(n, 9, 23)[6]
    assert int(_gpu_update(c, which, g, cut)) == _ref_update(sc, which, g, cut)
    rc = sc.state()
    assert (np.array_equal(c.xc(), rc["xc"]) if exact else rel_err(c.xc(), rc["xc"]).max() <= REL_TOL)
    # Arr::dot_mv over the group: all n entries
    x = np.random.default_rng(1).uniform(-1, 1, n)
    y = np.zeros(n)
    m = np.ascontiguousarray(r["mq"])
    ref.lib().ellref_dot_mv(m.ctypes.data, x.ctypes.data, n, n, y.ctypes.data)
    yg = e.dot_mv(x)
    assert yg.shape == (n,)
    assert np.array_equal(yg, y) if exact else np.abs(yg - y).max() <= REL_TOL * np.abs(y).max()
    e.set_xc(np.arange(n, dtype=float))
    assert np.array_equal(e.xc(), np.arange(n, dtype=float))


def test_handles_on_two_devices_in_one_process(ref):
    """the sweep's > 48 KB dynamic shared memory opt-in is per device"""
    if _ndev() < 2:
        pytest.skip("needs 2 GPUs in this process")
    n = 500
    kw = _start(n, "matrix", 3)
    es = [Ell(kw["kappa"], kw["xc"], mq=kw["mq"], device=d) for d in (0, 1)]
    s = ref.Space(A.KIND_ELL, n, kappa=kw["kappa"], mq=kw["mq"], xc=kw["xc"])
    for which, g, cut in _cuts(n, 5, 1):
        st_r = _ref_update(s, which, g, cut)
        for e in es:
            assert int(_gpu_update(e, which, g, cut)) == st_r
    r = s.state()
    for e in es:
        assert np.array_equal(e.xc(), r["xc"]) and np.array_equal(e.mq, r["mq"])
    x = np.random.default_rng(0).uniform(-1, 1, n)
    assert np.array_equal(es[1].dot_mv(x), es[0].dot_mv(x))      # dot_mv stages on the handle's device


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0, 0, 0], [0, 1], [0, 1, 2, 3]])
def test_batch_solve_multi(ref, devices):
    if max(devices) >= _ndev():
        pytest.skip("needs more GPUs in this process")
    for prob in (W.bowl(1001, n=32, max_iters=300), W.profit(3000), W.lmi(130, n=64, m=16),
                 W.lowpass(9, batch.lowpass_table(32), max_iters=500)):
        whole = batch.solve_batch(prob)
        multi = batch.solve_batch_multi(prob, devices)
        for f in ("num_iters", "exit_status", "has_xbest", "xbest", "gamma", "tsq", "xc", "kappa"):
            assert np.array_equal(getattr(multi, f), getattr(whole, f), equal_nan=True), f
        assert multi.launches == len(devices)
    empty = batch.solve_batch_multi(W.bowl(0, n=32), devices)
    assert empty.num_iters.shape == (0,)
    small = batch.solve_batch_multi(W.bowl(2, n=32, max_iters=50), devices)   # fewer problems than devices
    assert np.array_equal(small.num_iters, batch.solve_batch(W.bowl(2, n=32, max_iters=50)).num_iters)


def test_batch_solve_multi_bad_arguments():
    prob = W.bowl(4, n=32, max_iters=10)
    d = prob.to_desc(0)
    out = batch.BatchOutput.empty(4, 32).to_result()
    devs = (C.c_int32 * 1)(99)
    assert lib().ellcuda_batch_solve_multi(C.byref(d), C.byref(out), 1, devs) == A.E_ARG
    assert b"device" in lib().ellcuda_last_error()
    assert lib().ellcuda_batch_solve_multi(C.byref(d), C.byref(out), 0, devs) == A.E_ARG


# ---------------------------------------------------------------- symmetric storage (lower block-triangle only), 1e-9 gate
@pytest.mark.parametrize("n,start,mode", [(64, "scalar", 0), (100, "matrix", 0), (130, "diag", 3), (257, "matrix", 0), (700, "matrix", 2),
                                          (1024, "scalar", 0), (33, "matrix", 1)])
@pytest.mark.parametrize("no_defer", [False, True])
def test_symmetric_storage_trajectory(ref, n, start, mode, no_defer):
    kw = _start(n, start, n + 5)
    e = Ell(kw["kappa"], kw["xc"], mq=kw.get("mq"), diag=kw.get("diag"), no_defer_trick=no_defer, storage=A.STORAGE_SYMMETRIC,
            reduce_mode=mode)
    s = ref.Space(A.KIND_ELL, n, kappa=kw["kappa"], mq=kw.get("mq"), diag=kw.get("diag"), xc=kw["xc"], no_defer_trick=no_defer)
    r = s.state()
    assert np.array_equal(e.mq, r["mq"])                  # packing / unpacking is exact
    for k, (which, g, cut) in enumerate(_cuts(n, 12, n)):
        st_g, st_r = _gpu_update(e, which, g, cut), _ref_update(s, which, g, cut)
        r = s.state()
        assert int(st_g) == st_r, f"step {k}: status"
        assert rel_err(e.tsq(), r["tsq"]).max() <= REL_TOL and rel_err(e.kappa, r["kappa"]).max() <= REL_TOL, f"step {k}"
        assert np.abs(e.xc() - r["xc"]).max() <= REL_TOL * np.abs(r["xc"]).max(), f"step {k}: xc"
        if k in (3, 11):
            assert np.abs(e.mq - r["mq"]).max() <= REL_TOL * np.abs(r["mq"]).max(), f"step {k}: mq"
        if k == 5:
            c, sc = e.clone(), s.clone()
    which, g, cut = _cuts(n, 12, n)[6]
    assert int(_gpu_update(c, which, g, cut)) == _ref_update(sc, which, g, cut)
    assert np.abs(c.xc() - sc.state()["xc"]).max() <= REL_TOL * np.abs(sc.state()["xc"]).max()
    x = np.random.default_rng(2).uniform(-1, 1, n)        # Arr::dot_mv on the packed matrix
    y = np.zeros(n)
    m = np.ascontiguousarray(r["mq"])
    ref.lib().ellref_dot_mv(m.ctypes.data, x.ctypes.data, n, n, y.ctypes.data)
    assert np.abs(e.dot_mv(x) - y).max() <= REL_TOL * np.abs(y).max()


def test_symmetric_storage_is_deterministic_and_rejects_what_it_cannot_do():
    n = 300
    kw = _start(n, "matrix", 1)
    runs = []
    for _ in range(2):
        e = Ell(kw["kappa"], kw["xc"], mq=kw["mq"], storage=A.STORAGE_SYMMETRIC)
        for which, g, cut in _cuts(n, 6, 3):
            _gpu_update(e, which, g, cut)
        runs.append((e.tsq(), e.kappa, e.xc(), e.mq))
    assert runs[0][0] == runs[1][0] and runs[0][1] == runs[1][1]
    assert np.array_equal(runs[0][2], runs[1][2]) and np.array_equal(runs[0][3], runs[1][3])
    with pytest.raises(EllCudaError):
        Ell(1.0, np.zeros(4), mq=_asym(4, 0), storage=A.STORAGE_SYMMETRIC)       # needs a symmetric matrix
    with pytest.raises(EllCudaError):
        Ell(1.0, np.zeros(128), rows=(0, 64), storage=A.STORAGE_SYMMETRIC)        # does not shard rows


def test_symmetric_storage_drives_the_reference_loop():  # src/example1.rs:39-70 → 25 iterations, through the packed matrix
    from ellalgo_rs_b200.cutting_plane import Options, cutting_plane_optim
    from test_stream_gpu import _Example1
    e = Ell.new_with_scalar(10.0, np.zeros(2), storage=A.STORAGE_SYMMETRIC)
    xbest, niter, gamma = cutting_plane_optim(_Example1(), e, -math.inf, Options(tolerance=1e-10))
    assert xbest is not None and niter == 25


@pytest.mark.parametrize("devices", [[0, 0], [0, 0, 0], [0, 1], [0, 1, 2, 3]])
@pytest.mark.parametrize("n", [130, 1024])
def test_symmetric_storage_group(ref, n, devices):
    """symmetric storage sharded inside one process: column-block pairs dealt round-robin, the partial grad_t reduced over the slabs"""
    if max(devices) >= _ndev():
        pytest.skip("needs more GPUs in this process")
    kw = _start(n, "matrix", 41)
    e = Ell(kw["kappa"], kw["xc"], mq=kw["mq"], devices=devices, storage=A.STORAGE_SYMMETRIC)
    s = ref.Space(A.KIND_ELL, n, kappa=kw["kappa"], mq=kw["mq"], xc=kw["xc"])
    assert np.array_equal(e.mq, kw["mq"])
    for k, (which, g, cut) in enumerate(_cuts(n, 10, 29)):
        st_g, st_r = _gpu_update(e, which, g, cut), _ref_update(s, which, g, cut)
        r = s.state()
        assert int(st_g) == st_r, f"step {k}"
        assert rel_err(e.tsq(), r["tsq"]).max() <= REL_TOL and rel_err(e.kappa, r["kappa"]).max() <= REL_TOL, f"step {k}"
        assert np.abs(e.xc() - r["xc"]).max() <= REL_TOL * np.abs(r["xc"]).max(), f"step {k}"
        if k == 6:
            c, sc = e.clone(), s.clone()
    assert np.abs(e.mq - r["mq"]).max() <= REL_TOL * np.abs(r["mq"]).max()
    which, g, cut = _cuts(n, 10, 29)[7]
    assert int(_gpu_update(c, which, g, cut)) == _ref_update(sc, which, g, cut)
    assert np.abs(c.xc() - sc.state()["xc"]).max() <= REL_TOL * np.abs(sc.state()["xc"]).max()
