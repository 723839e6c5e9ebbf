"""tiny run of every kernel family, for compute-sanitizer (memcheck / racecheck); checks results against nothing —
the sanitizer's report is the output"""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ellalgo_rs_b200 import batch, workloads as W, _abi as A
from ellalgo_rs_b200.problem import BatchProblem
from ellalgo_rs_b200.ell import Ell, EllStable
from ellalgo_rs_b200.cutting_plane import ParallelCut, SingleCut

o = batch.solve_batch(W.example1(3)); print("example1", o.num_iters)
o = batch.solve_batch(W.profit(70)); print("profit", o.num_iters[:3])
o = batch.solve_batch(W.bowl(20, n=32, max_iters=60)); print("bowl32", o.num_iters[:3])
o = batch.solve_batch(W.bowl(9, n=7, max_iters=60)); print("bowl7", o.num_iters[:3])
t = batch.lowpass_table(32)
o = batch.solve_batch(W.lowpass(10, t, max_iters=150)); print("lowpass", o.num_iters[:3])
o = batch.solve_batch(W.lmi(40, n=64, m=16)); print("lmi diag", o.num_iters[:3])
o = batch.solve_batch(W.lmi(10, n=64, m=16, general=True, upper_scale=1e-6)); print("lmi general", o.num_iters[:3])
p = BatchProblem(oracle_id=A.ORACLE_EXAMPLE3, n=2, count=5, mode=A.MODE_BSEARCH, kappa0=100.0, gamma0=-100.0, params=[100.0, 1e-8, 2000.0], tolerance=1e-8)
o = batch.solve_batch(p); print("bsearch", o.num_iters)
for n, mode in ((200, 0), (130, 3)):
    e = Ell.new_with_scalar(1.0, np.zeros(n), reduce_mode=mode)
    for k in range(4):
        e.update_bias_cut((W.large_n_grad(n, k), ParallelCut(*W.LARGE_N_CUT)))
    print("ell", n, mode, e.kappa, e.mq.shape)
parts = [Ell.new_with_scalar(1.0, np.zeros(128), rows=(64 * p, 64 * (p + 1)), reduce_mode=2) for p in range(2)]
st = torch.cuda.Stream()
parts = [Ell.new_with_scalar(1.0, np.zeros(128), rows=(64 * p, 64 * (p + 1)), reduce_mode=2, stream=st.cuda_stream) for p in range(2)]
blocks = [q.exchange_block_ptr() for q in parts]
for p, q in enumerate(parts):
    q.exchange_attach_ptrs(blocks, p)
with torch.cuda.stream(st):
    for k in range(3):
        g = torch.from_numpy(W.large_n_grad(128, k)).cuda()
        for q in parts: q.update_begin(A.UPDATE_BIAS, g.data_ptr(), ParallelCut(*W.LARGE_N_CUT))
        for q in parts: q.update_finish()
    print("peer exchange", [int(q.last_status()) for q in parts])
s = EllStable.new_with_scalar(2.0, np.zeros(40))
for k in range(3):
    s.update_bias_cut((W.large_n_grad(40, k), SingleCut(0.01)))
print("ellstable", s.kappa)
