"""kernel-only timing of the n=2 thread-per-problem kernel on 2^20 profit problems (device-resident inputs)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ellalgo_rs_b200 import batch, workloads as W
from ellalgo_rs_b200.problem import BatchProblem
cnt = 1 << 20
h = W.profit(cnt)
kw = {k: torch.from_numpy(np.ascontiguousarray(getattr(h, k))).cuda() for k in ("kappa0", "xc0", "diag0", "gamma0", "params")}
p = BatchProblem(oracle_id=h.oracle_id, n=2, count=cnt, max_iters=h.max_iters, tolerance=h.tolerance, **kw)
res = {"num_iters": torch.zeros(cnt, dtype=torch.int64, device="cuda"), "exit_status": torch.zeros(cnt, dtype=torch.int32, device="cuda"),
       "xbest": torch.zeros(cnt, 2, dtype=torch.float64, device="cuda"), "has_xbest": torch.zeros(cnt, dtype=torch.uint8, device="cuda"),
       "gamma": torch.zeros(cnt, dtype=torch.float64, device="cuda")}
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
best = 1e9
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s); batch.solve_batch_device(p, res, s.cuda_stream); e1.record(s); e1.synchronize()
    best = min(best, e0.elapsed_time(e1))
it = int(res["num_iters"].sum()) + cnt
print(f"profit 2^20: kernel {best:.2f} ms, {it/best/1e3:.0f} M iter/s, {cnt/best/1e3:.1f} M problems/s", flush=True)
