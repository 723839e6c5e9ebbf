"""a P-way row-sharded n=16384 update emulated on ONE GPU: P slab handles attached to each other by pointer, all on one
stream (begin for every rank, then finish for every rank — no kernel ever waits for a later launch).  The P sweeps run back
to back, so (time per update) / P is one rank's share of the work, without the cross-GPU overlap.  Eager and CUDA-graph
per-update times on the device; run under ncu for per-kernel durations.   usage: profile_slab.py [P] [updates] [eager-only]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ellalgo_rs_b200 import workloads as W, _abi as A
from ellalgo_rs_b200.ell import Ell
from ellalgo_rs_b200.cutting_plane import ParallelCut
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
NSEG = int(os.environ.get('SLAB_NSEG', P))
n = int(os.environ.get('SLAB_N', 16384))
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
gs = [torch.from_numpy(W.large_n_grad(n, k)).cuda() for k in range(20)]
cut = ParallelCut(*W.LARGE_N_CUT)

def fresh():
    b = [n * r // P for r in range(P + 1)]
    es = [Ell.new_with_scalar(1.0, np.zeros(n), rows=(b[r], b[r + 1]), reduce_mode=NSEG, stream=s.cuda_stream) for r in range(P)]
    blocks = [e.exchange_block_ptr() for e in es]
    for r, e in enumerate(es):
        e.exchange_attach_ptrs(blocks, r)
    for k in range(10):
        step(es, k)
    es[0].sync()
    return es

def step(es, k):
    for e in es:
        e.update_begin(A.UPDATE_BIAS, gs[k % 20].data_ptr(), cut)
    for e in es:
        e.update_finish()

es = fresh()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for k in range(K):
    step(es, k)
e1.record(s); e1.synchronize()
t = e0.elapsed_time(e1) / K * 1e3
print(f"P={P} nseg={NSEG}: {t:.1f} us per update eager = {t / P:.1f} us per rank (status {es[0].last_status()})", flush=True)
if len(sys.argv) > 3:
    sys.exit(0)
del es
es = fresh()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    for k in range(20):
        step(es, k)
g.replay(); torch.cuda.synchronize()
e0.record(s)
for _ in range(5):
    g.replay()
e1.record(s); e1.synchronize()
t = e0.elapsed_time(e1) / 100 * 1e3
print(f"P={P} nseg={NSEG}: {t:.1f} us per update as CUDA-graph replays = {t / P:.1f} us per rank (status {es[0].last_status()}, kappa {es[0].kappa:.6g})", flush=True)
