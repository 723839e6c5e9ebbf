"""N>1 host logic on CPU: world_size-2 gloo.  The per-rank compute is stood in for by the CPU checker (allowed in
tests); what is under test is the sharding, the exchange pattern and the gather of ellalgo_rs_b200/parallel.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _init(rank, world, port):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _batch_worker(rank, world, port, q):
    _init(rank, world, port)
    import ellref_py as R
    from ellalgo_rs_b200 import workloads as W
    from ellalgo_rs_b200.parallel import solve_batch_sharded
    prob = W.profit(101)                       # odd count: uneven shards
    out = solve_batch_sharded(prob, lambda p: R.solve(p, 1), rank, world)
    if rank == 0:
        whole = R.solve(prob, 1)
        ok = all(np.array_equal(getattr(out, f), getattr(whole, f)) for f in ("num_iters", "exit_status", "xbest", "gamma", "has_xbest"))
        q.put(bool(ok) and out.num_iters[0] == 83)
    else:
        assert out is None
    dist.destroy_process_group()


def _slab_worker(rank, world, port, q):
    """row-sharded Ell update, emulated with numpy per rank: slab GEMV -> all_gather(grad_t) -> redundant finish."""
    _init(rank, world, port)
    import ellref_py as R
    from ellalgo_rs_b200 import _abi as A
    from ellalgo_rs_b200.parallel import shard_range
    n = 24
    rng = np.random.default_rng(5)
    a = rng.uniform(-1, 1, (n, n)) / n
    mq = np.eye(n) + a + a.T
    xc = rng.uniform(-1, 1, n)
    ref = R.Space(A.KIND_ELL, n, kappa=2.0, mq=mq, xc=xc)
    lo, hi = shard_range(n, rank, world)
    slab = mq[:, lo:hi].copy()                 # column block = this rank's rows, M being symmetric
    kappa, x = 2.0, xc.copy()
    ok = True
    for k in range(5):
        g = rng.uniform(-1, 1, n)
        # local part of grad_t: sum over i ascending, per owned column (the device kernel's order)
        local = np.zeros(hi - lo)
        for c in range(hi - lo):
            acc = 0.0
            for i in range(n):
                acc += slab[i, c] * g[i]
            local[c] = acc
        parts = [torch.zeros(shard_range(n, r, world)[1] - shard_range(n, r, world)[0], dtype=torch.float64) for r in range(world)]
        dist.all_gather(parts, torch.from_numpy(local)) if n % world == 0 else None
        gt = torch.cat(parts).numpy()
        omega = 0.0
        for i in range(n):
            omega += g[i] * gt[i]
        tsq = kappa * omega
        st, (rho, sigma, delta) = R.calc(n, 0, 0.01, 0.10, tsq)
        st_ref = ref.update(0, g, 0.01, 0.10, single=False)
        ok &= st == st_ref == 0
        x = x - (rho / omega) * gt
        ratio = sigma / omega
        for c in range(hi - lo):
            cg = lo + c
            for i in range(n):
                h, l = max(i, cg), min(i, cg)
                slab[i, c] -= (ratio * gt[h]) * gt[l]
        kappa *= delta
        s = ref.state()
        ok &= np.array_equal(x, s["xc"]) and kappa == s["kappa"] and tsq == s["tsq"]
        ok &= np.array_equal(slab, s["mq"][:, lo:hi])
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    if rank == 0:
        q.put(all(flags))
    dist.destroy_process_group()


def _run(worker, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return q.get(timeout=5)


def test_shard_range_partitions():
    from ellalgo_rs_b200.parallel import shard_range
    for total in (0, 1, 7, 64, 65536, 101):
        for world in (1, 2, 3, 8):
            parts = [shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_batch_sharded_over_two_ranks_gloo(ref):
    assert _run(_batch_worker, 29611)


def test_row_sharded_update_over_two_ranks_gloo(ref):
    assert _run(_slab_worker, 29612)
