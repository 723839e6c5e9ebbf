"""a P-way sharded symmetric-storage n=16384 update emulated on ONE GPU (ellcuda_create_sharded with the same device P times:
the slabs share one stream, so the P sweeps run back to back): (time per update) / P = one rank's share of the device work.
usage: profile_symslab.py [P] [nseg]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ellalgo_rs_b200 import workloads as W, _abi as A
from ellalgo_rs_b200.ell import Ell
from ellalgo_rs_b200.cutting_plane import ParallelCut
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nseg = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(os.environ.get("SLAB_N", 16384))
gs = [torch.from_numpy(W.large_n_grad(n, k)).cuda() for k in range(20)]
cut = ParallelCut(*W.LARGE_N_CUT)
e = Ell.new_with_scalar(1.0, np.zeros(n), devices=[0] * P, storage=A.STORAGE_SYMMETRIC, reduce_mode=nseg)
for k in range(10):
    e.update_begin(A.UPDATE_BIAS, gs[k % 20].data_ptr(), cut); e.update_finish()
e.sync()
K = int(os.environ.get('SLAB_K', 100))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for k in range(K):
    e.update_begin(A.UPDATE_BIAS, gs[k % 20].data_ptr(), cut); e.update_finish()
e.sync()
t = (time.perf_counter() - t0) / K * 1e6
print(f"P={P} nseg={nseg or 'auto'}: {t:.1f} us per update (host clock, eager) = {t / P:.1f} us per rank (status {e.last_status()}, kappa {e.kappa:.6g})", flush=True)
