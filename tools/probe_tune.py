"""times the n=32 batched kernels for the current ELLCUDA_TUNE variant (device events, device-resident inputs)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ellalgo_rs_b200 import batch, workloads as W, _abi as A
from ellalgo_rs_b200.problem import BatchProblem

def dev_problem(host, keys):
    kw = {k: torch.from_numpy(np.ascontiguousarray(getattr(host, k))).cuda() for k in keys}
    return BatchProblem(oracle_id=host.oracle_id, n=host.n, count=host.count, max_iters=host.max_iters, tolerance=host.tolerance,
                        aux=list(host.aux), **kw)

def run(name, host, keys, reps=3):
    p = dev_problem(host, keys)
    res = {"num_iters": torch.zeros(host.count, dtype=torch.int64, device="cuda"),
           "exit_status": torch.zeros(host.count, dtype=torch.int32, device="cuda"),
           "xbest": torch.zeros(host.count, host.n, dtype=torch.float64, device="cuda"),
           "has_xbest": torch.zeros(host.count, dtype=torch.uint8, device="cuda"),
           "gamma": torch.zeros(host.count, dtype=torch.float64, device="cuda")}
    s = torch.cuda.Stream(); torch.cuda.set_stream(s)
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); batch.solve_batch_device(p, res, s.cuda_stream); e1.record(s); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    it = int(res["num_iters"].sum()) + int((res["exit_status"] != 5).sum())
    print(f"TUNE={os.environ.get('ELLCUDA_TUNE','0')} {name}: {it/best/1e3:.1f} M iter/s ({best:.1f} ms, checksum {int(res['num_iters'].sum())})", flush=True)

t = batch.lowpass_table(32)
run("lowpass32 x16384", W.lowpass(16384, t), ("kappa0", "xc0", "gamma0", "params", "table"))
run("bowl32 x65536 maxit2000", W.bowl(65536, n=32, max_iters=2000), ("kappa0", "xc0", "gamma0", "params"))
