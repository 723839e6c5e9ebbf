"""symmetric-storage large-n update: per-update time as CUDA-graph replays (ELLCUDA_SYM_TMA / SYM_NSEG select variants)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ellalgo_rs_b200 import workloads as W, _abi as A
from ellalgo_rs_b200.ell import Ell
from ellalgo_rs_b200.cutting_plane import ParallelCut
n = int(os.environ.get("SYM_N", 16384))
storage = int(os.environ.get("SYM_STORAGE", 1))
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
gs = [torch.from_numpy(W.large_n_grad(n, k)).cuda() for k in range(20)]
cut = ParallelCut(*W.LARGE_N_CUT)
e = Ell.new_with_scalar(1.0, np.zeros(n), storage=storage, reduce_mode=int(os.environ.get("SYM_NSEG", 0)), stream=s.cuda_stream)
for k in range(10):
    e.update_begin(A.UPDATE_BIAS, gs[k % 20].data_ptr(), cut); e.update_finish()
e.sync()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    for k in range(20):
        e.update_begin(A.UPDATE_BIAS, gs[k].data_ptr(), cut); e.update_finish()
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(5):
    g.replay()
e1.record(s); e1.synchronize()
t = e0.elapsed_time(e1) / 100 * 1e3
print(f"n={n} storage={storage} SYM_TMA={os.environ.get('ELLCUDA_SYM_TMA','1')} nseg={os.environ.get('SYM_NSEG','auto')}: {t:.1f} us per update = {1e6/t:.0f} updates/s "
      f"(status {e.last_status()}, kappa {e.kappa:.6g})", flush=True)
