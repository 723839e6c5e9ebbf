"""quick timings of the batched Ell kernels for n <= 64: LMI on Ell n=64 m=16 and lowpass n=40 run in the TMEM kernel; bowl n=48/64 and lowpass n=64 go to the CTA kernel unless ELLCUDA_CTA_MIN_N=65"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ellalgo_rs_b200 import batch, workloads as W, _abi as A

def run(name, p, reps=2):
    for _ in range(reps):
        t0 = time.perf_counter(); o = batch.solve_batch(p); dt = time.perf_counter() - t0
    it = int(o.num_iters.sum()) + int((o.exit_status != 5).sum())
    print(f"{name}: {it/dt/1e6:.1f} M iter/s ({dt*1e3:.1f} ms, host wall clock incl. copies, {it} iterations, checksum {int(o.num_iters.sum())})", flush=True)

run("bowl n=64 x9472 maxit2000", W.bowl(9472, n=64, max_iters=2000))
run("bowl n=48 x9472 maxit2000", W.bowl(9472, n=48, max_iters=2000))
run("bowl n=32 x18944 maxit2000 (registers kernel)", W.bowl(18944, n=32, max_iters=2000))
p = W.lmi(1 << 15, n=64, m=16); p.kind = A.KIND_ELL
run("lmi n=64 m=16 on Ell x32768", p)
run("lmi n=64 m=16 on EllStable x32768", W.lmi(1 << 15, n=64, m=16))
run("lowpass n=64 x2368 maxit3000", W.lowpass(2368, batch.lowpass_table(64), n=64, max_iters=3000))
run("lowpass n=40 x2368 maxit3000", W.lowpass(2368, batch.lowpass_table(40), n=40, max_iters=3000))
