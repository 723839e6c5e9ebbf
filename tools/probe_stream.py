"""rough first-look timing of the large-n sweep (not the bench)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ellalgo_rs_b200 import workloads as W, _abi as A
from ellalgo_rs_b200.ell import Ell
from ellalgo_rs_b200.cutting_plane import ParallelCut

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for mode in [int(a) for a in sys.argv[2:]] or [0]:
    e = Ell.new_with_scalar(1.0, np.zeros(n), reduce_mode=mode)
    gs = [torch.from_numpy(W.large_n_grad(n, k)).cuda() for k in range(8)]
    torch.cuda.synchronize()
    cut = ParallelCut(*W.LARGE_N_CUT)
    e.set_timing(True)
    sweeps = []
    for k in range(4):
        e.update_begin(A.UPDATE_BIAS, gs[k % 8].data_ptr(), cut); e.update_finish(); e.sync()
    t0 = time.perf_counter()
    K = 30
    for k in range(K):
        e.update_begin(A.UPDATE_BIAS, gs[k % 8].data_ptr(), cut); e.update_finish()
        sweeps.append(e.last_sweep_ms())
    e.sync()
    dt = (time.perf_counter() - t0) / K
    sw = float(np.median(sweeps))
    print(f"n={n} mode={mode}: {dt*1e3:.3f} ms/update (loop incl. event syncs), sweep median {sw:.3f} ms min {min(sweeps):.3f} "
          f"→ {16.0*n*n/sw/1e6:.0f} GB/s algorithmic; status={e.last_status()} kappa={e.kappa:.6g}", flush=True)
    del e

# host-gradient path (SearchSpace::update_bias_cut through ellcuda_update): H2D + 4 kernels + status read-back per call
e = Ell.new_with_scalar(1.0, np.zeros(n))
hg = [torch.from_numpy(W.large_n_grad(n, k)).pin_memory().numpy() for k in range(8)]
for k in range(5):
    e.update_bias_cut((hg[k % 8], cut))
t0 = time.perf_counter()
for k in range(30):
    e.update_bias_cut((hg[k % 8], cut))
dt = (time.perf_counter() - t0) / 30
print(f"n={n} host-gradient update: {dt*1e3:.3f} ms/update → {16.0*n*n/dt/1e9:.0f} GB/s algorithmic", flush=True)
