"""rough first-look timings of the batched kernels (not the bench)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ellalgo_rs_b200 import batch, workloads as W

def run(name, prob, reps=2):
    for _ in range(reps):
        t0 = time.perf_counter(); o = batch.solve_batch(prob); dt = time.perf_counter() - t0
    it = int(o.num_iters.sum()) + int((o.exit_status != 5).sum())  # updates performed ≈ num_iters (+1 for early exits)
    print(f"{name}: count={prob.count} total_iters={it} time={dt*1e3:.1f} ms  {it/dt/1e6:.1f} M iter/s  "
          f"mean_iters={o.num_iters.mean():.1f} exits={np.bincount(o.exit_status, minlength=7).tolist()}", flush=True)

run("bowl32 maxit2000", W.bowl(1 << 16, n=32, max_iters=2000))
run("bowl32 full", W.bowl(1 << 14, n=32, max_iters=20000))
t = batch.lowpass_table(32)
run("lowpass32", W.lowpass(4096, t))
run("profit", W.profit(1 << 20))
run("lmi64 diag", W.lmi(1 << 15, n=64, m=16))
run("lmi64 general 1e-6", W.lmi(1 << 12, n=64, m=16, general=True, upper_scale=1e-6))
