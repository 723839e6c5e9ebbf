import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ellalgo_rs_b200 import batch, workloads as W
p = W.lmi(1 << 15, n=64, m=16)
for _ in range(2):
    t0 = time.perf_counter(); o = batch.solve_batch(p); dt = time.perf_counter() - t0
it = int(o.num_iters.sum()) + int((o.exit_status != 5).sum())
print(f"LMI_WARPS={os.environ.get('ELLCUDA_LMI_WARPS','16')}: lmi64 diag {it/dt/1e6:.1f} M iter/s ({dt*1e3:.1f} ms) checksum {int(o.num_iters.sum())}", flush=True)
