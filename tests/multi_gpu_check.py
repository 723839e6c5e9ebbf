"""torchrun --nproc-per-node P tests/multi_gpu_check.py — NCCL row-sharded Ell against the CPU checker (bit for bit in
reduce_mode 0, 1e-9 in reduce_mode 4) and problem-sharded batch against the single-rank result."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, torch, torch.distributed as dist
import ellref_py as R
from ellalgo_rs_b200 import _abi as A, batch, workloads as W
from ellalgo_rs_b200.cutting_plane import ParallelCut
from ellalgo_rs_b200.parallel import ShardedEll, solve_batch_sharded

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
ok = True
for n, mode, xch in ((1024, 0, 'p2p'), (1000, 0, 'nccl'), (2048, 4, 'p2p'), (4096, 0, 'p2p')):
    rng = np.random.default_rng(3)
    a = rng.uniform(-1, 1, (n, n)) / n
    mq = np.eye(n) + a + a.T
    xc = rng.uniform(-1, 1, n)
    s = R.Space(A.KIND_ELL, n, kappa=1.0, mq=mq, xc=xc)
    e = ShardedEll(1.0, xc, mq=mq, reduce_mode=mode, device=local, exchange=xch)
    for k in range(6):
        g = W.large_n_grad(n, k)
        beta = ParallelCut(0.01, 0.10) if k != 3 else ParallelCut(5.0, 4.0)   # k=3: NoSoln, state must not move
        st = e.update(A.UPDATE_BIAS, torch.from_numpy(g).cuda(), beta)
        st_r = s.update(0, g, beta.beta0, beta.beta1, single=False)
        r = s.state()
        if mode == 0:
            good = int(st) == st_r and e.tsq() == r["tsq"] and e.kappa == r["kappa"] and np.array_equal(e.xc(), r["xc"])
        else:
            good = int(st) == st_r and abs(e.tsq() - r["tsq"]) <= 1e-9 * abs(r["tsq"]) and np.abs(e.xc() - r["xc"]).max() <= 1e-9 * np.abs(r["xc"]).max()
        ok &= bool(good)
    slab = e.ell.mq
    r = s.state()
    ok &= bool(np.array_equal(slab, r["mq"][e.lo:e.hi]) if mode == 0 else np.abs(slab - r["mq"][e.lo:e.hi]).max() <= 1e-9)
    print(f"rank {rank}: n={n} mode={mode} exchange={xch} rows [{e.lo},{e.hi}) ok={ok}", flush=True)
# symmetric storage sharded over the ranks (column-block pairs round-robin, reduction of flagged lines): 1e-9 gate, and each
# rank's OWN tiles of M against the checker
for n in (1024, 700):
    rng = np.random.default_rng(5)
    a = rng.uniform(-1, 1, (n, n)) / n
    mq = np.eye(n) + a + a.T
    xc = rng.uniform(-1, 1, n)
    s = R.Space(A.KIND_ELL, n, kappa=1.0, mq=mq, xc=xc)
    e = ShardedEll(1.0, xc, mq=mq, device=local, storage=A.STORAGE_SYMMETRIC)
    good = True
    for k in range(7):
        g = W.large_n_grad(n, k)
        beta = ParallelCut(0.01, 0.10) if k != 3 else ParallelCut(5.0, 4.0)
        st = e.update(A.UPDATE_BIAS, torch.from_numpy(g).cuda(), beta)
        st_r = s.update(0, g, beta.beta0, beta.beta1, single=False)
        r = s.state()
        good &= int(st) == st_r and abs(e.tsq() - r["tsq"]) <= 1e-9 * abs(r["tsq"]) and np.abs(e.xc() - r["xc"]).max() <= 1e-9 * np.abs(r["xc"]).max()
    T = (n + 63) // 64
    blk = np.arange(n) // 64
    J = np.minimum(blk[:, None], blk[None, :])
    mine = (np.minimum(J, T - 1 - J) % world) == rank
    full, r = e.ell.mq, s.state()
    good &= bool(np.abs(full - r["mq"])[mine].max(initial=0.0) <= 1e-9 * np.abs(r["mq"]).max())   # a rank may own no pair
    ok &= bool(good)
    print(f"rank {rank}: n={n} symmetric storage sharded over {world} ranks ok={good}", flush=True)
    del e
# the configuration bench.py times at this world size: n = 16384, reduce_mode = world, p2p — first two updates from M = I
# against the checker (rank 0 runs it: ~15 s per update on one core), every rank compares its scalars, xc and its slab of M
if os.environ.get("ELL_MGC_LARGE", "1") != "0":
    import ctypes as C
    n, mode = 16384, world
    e = ShardedEll(1.0, np.zeros(n), reduce_mode=mode, device=local, exchange="p2p")
    got = []
    for k in range(2):
        st = e.update(A.UPDATE_BIAS, torch.from_numpy(W.large_n_grad(n, k)).cuda(), ParallelCut(*W.LARGE_N_CUT))
        got.append((int(st), e.tsq(), e.kappa, e.xc()))
    slab = e.ell.mq
    ref = [None]
    if rank == 0:
        mq, xc = np.eye(n), np.zeros(n)
        kap, tsq = C.c_double(1.0), C.c_double(0.0)
        states = []
        for k in range(2):
            g = W.large_n_grad(n, k)
            st = R.lib().ellref_ell_update_inplace(n, mq.ctypes.data, xc.ctypes.data, C.byref(kap), C.byref(tsq), 0, 1, 0, g.ctypes.data,
                                                   W.LARGE_N_CUT[0], 1, W.LARGE_N_CUT[1], None)
            states.append((int(st), tsq.value, kap.value, xc.copy()))
        bounds = [n * r // world for r in range(world + 1)]
        ref = [(states, [mq[bounds[r]:bounds[r + 1]] for r in range(world)])]
    dist.broadcast_object_list(ref, src=0)
    states, slabs = ref[0]
    worst = 0.0
    for (st, t, kp, x), (st_r, t_r, kp_r, x_r) in zip(got, states):
        ok &= st == st_r
        worst = max(worst, abs(t - t_r) / abs(t_r), abs(kp - kp_r) / abs(kp_r), float(np.abs(x - x_r).max() / np.abs(x_r).max()))
    worst = max(worst, float(np.abs(slab - slabs[rank]).max() / np.abs(slabs[rank]).max()))
    ok &= worst <= 1e-9
    print(f"rank {rank}: n={n} mode={mode} exchange=p2p (the bench configuration) max rel err {worst:.3g} ok={ok}", flush=True)
    del e, slab, ref, slabs
prob = W.bowl(1001, n=32, max_iters=300)
prob.device = local
out = solve_batch_sharded(prob, batch.solve_batch, rank, world)
if rank == 0:
    whole = batch.solve_batch(prob)
    ok &= all(np.array_equal(getattr(out, f), getattr(whole, f)) for f in ("num_iters", "exit_status", "xbest", "gamma"))
flags = [None] * world
dist.all_gather_object(flags, bool(ok))
if rank == 0:
    print("SHARDED CHECK", "PASSED" if all(flags) else "FAILED", flags, flush=True)
dist.destroy_process_group()
sys.exit(0 if all(flags) else 1)
