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
