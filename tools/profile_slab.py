"""one rank's work of a P-way row-sharded n=16384 update, on one GPU (peer exchange attached to itself only):
eager vs CUDA-graph per-update time on the device; run under ncu for per-kernel durations.
(The other ranks' slabs of grad_t never arrive, so omega is ~1/P of the real one: the cut is scaled to stay valid.)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ellalgo_rs_b200 import workloads as W, _abi as A
from ellalgo_rs_b200.ell import Ell
from ellalgo_rs_b200.cutting_plane import ParallelCut
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
NSEG = int(os.environ.get('SLAB_NSEG', P))
n = 16384
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
gs = [torch.from_numpy(W.large_n_grad(n, k)).cuda() for k in range(20)]
cut = ParallelCut(0.01 / P, 0.10 / P)

def fresh():
    e = Ell.new_with_scalar(1.0, np.zeros(n), rows=(0, n // P), reduce_mode=NSEG, stream=s.cuda_stream)
    e.exchange_attach_ptrs([e.exchange_block_ptr()], 0)
    for k in range(10):
        e.update_begin(A.UPDATE_BIAS, gs[k % 20].data_ptr(), cut); e.update_finish()
    e.sync()
    return e

e = fresh()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for k in range(K):
    e.update_begin(A.UPDATE_BIAS, gs[k % 20].data_ptr(), cut); e.update_finish()
e1.record(s); e1.synchronize()
print(f"P={P}: {e0.elapsed_time(e1)/K*1e3:.1f} us per update eager (status {e.last_status()})", flush=True)
if len(sys.argv) > 3:
    sys.exit(0)
e = fresh()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    for k in range(20):
        e.update_begin(A.UPDATE_BIAS, gs[k].data_ptr(), cut); e.update_finish()
g.replay(); torch.cuda.synchronize()
e0.record(s)
for _ in range(5):
    g.replay()
e1.record(s); e1.synchronize()
print(f"P={P} nseg={NSEG}: {e0.elapsed_time(e1)/100*1e3:.1f} us per update as CUDA-graph replays (status {e.last_status()})", flush=True)
