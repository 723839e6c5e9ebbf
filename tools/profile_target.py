"""short, deterministic workload for ncu: one lowpass n=32 batch launch + a few large-n updates"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ellalgo_rs_b200 import batch, workloads as W, _abi as A
from ellalgo_rs_b200.ell import Ell
from ellalgo_rs_b200.cutting_plane import ParallelCut

what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "batch"):
    t = batch.lowpass_table(32)
    o = batch.solve_batch(W.lowpass(int(os.environ.get("PROF_COUNT", "4736")), t))   # 2 problems per resident warp
    print("lowpass iterations", int(o.num_iters.sum()))
    o = batch.solve_batch(W.bowl(int(os.environ.get("PROF_COUNT", "4736")) * 2, n=32, max_iters=2000))
    print("bowl iterations", int(o.num_iters.sum()))
if what in ("all", "stream"):
    n = 16384
    e = Ell.new_with_scalar(1.0, np.zeros(n))
    gs = [torch.from_numpy(W.large_n_grad(n, k)).cuda() for k in range(4)]
    for k in range(5):
        e.update_begin(A.UPDATE_BIAS, gs[k % 4].data_ptr(), ParallelCut(*W.LARGE_N_CUT)); e.update_finish()
    e.sync()
    print("large-n status", e.last_status(), e.kappa)

if what == "headline":   # the bench's launch shape: 65536 lowpass problems, one launch
    t = batch.lowpass_table(32)
    o = batch.solve_batch(W.lowpass(65536, t))
    print("lowpass iterations", int(o.num_iters.sum()) + int((o.exit_status != 5).sum()))
if what == "sym":        # symmetric storage, n = 16384: a few updates
    n = 16384
    e = Ell.new_with_scalar(1.0, np.zeros(n), storage=A.STORAGE_SYMMETRIC)
    gs = [torch.from_numpy(W.large_n_grad(n, k)).cuda() for k in range(4)]
    for k in range(5):
        e.update_begin(A.UPDATE_BIAS, gs[k % 4].data_ptr(), ParallelCut(*W.LARGE_N_CUT)); e.update_finish()
    e.sync()
    print("sym status", e.last_status(), e.kappa)
if what == "lmi":
    o = batch.solve_batch(W.lmi(1 << 15, n=64, m=16))
    print("lmi iterations", int(o.num_iters.sum()) + int((o.exit_status != 5).sum()))
if what == "cta":        # one CTA per problem, packed triangle of M in shared memory: one wave of bowl and of lowpass at n = 128
    o = batch.solve_batch(W.bowl(444, n=128, max_iters=500))
    print("bowl128 iterations", int(o.num_iters.sum()) + int((o.exit_status != 5).sum()))
    o = batch.solve_batch(W.lowpass(444, batch.lowpass_table(128), n=128, max_iters=300))
    print("lowpass128 iterations", int(o.num_iters.sum()) + int((o.exit_status != 5).sum()))
