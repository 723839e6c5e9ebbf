"""quick timings of the batched Ell kernel for 64 < n <= 224 (M in shared memory, one CTA per problem): bowl and lowpass"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ellalgo_rs_b200 import batch, workloads as W, _abi as A

def run(name, p, reps=2):
    for _ in range(reps):
        t0 = time.perf_counter(); o = batch.solve_batch(p); dt = time.perf_counter() - t0
    it = int(o.num_iters.sum()) + int((o.exit_status != 5).sum())
    n = p.n
    print(f"{name}: {it/dt/1e6:.2f} M iter/s = {it/dt*16*n*n/1e12:.2f} TB/s equivalent streaming ({dt*1e3:.1f} ms host wall clock incl. copies, "
          f"{it} iterations, checksum {int(o.num_iters.sum())})", flush=True)

run("bowl n=128 x3552 maxit2000", W.bowl(3552, n=128, max_iters=2000))
run("bowl n=96 x3552 maxit2000", W.bowl(3552, n=96, max_iters=2000))
run("bowl n=160 x1184 maxit1000", W.bowl(1184, n=160, max_iters=1000))
run("bowl n=224 x592 maxit1000", W.bowl(592, n=224, max_iters=1000))
run("bowl n=64 x9472 maxit2000 (TMEM kernel, for comparison)", W.bowl(9472, n=64, max_iters=2000))
run("lowpass n=128 x1776 maxit3000", W.lowpass(1776, batch.lowpass_table(128), n=128, max_iters=3000))
run("lowpass n=96 x888 maxit3000", W.lowpass(888, batch.lowpass_table(96), n=96, max_iters=3000))
