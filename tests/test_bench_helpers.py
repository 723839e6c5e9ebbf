"""CPU checks of bench.py's host-side helpers: the strided sub-batch used for the CPU baseline / parity spot check is the
same problems as the big batch (so comparing sample[k] with gpu[k*stride] is meaningful), and the iteration count convention."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ellalgo_rs_b200 import workloads as W  # noqa: E402


def test_strided_subbatch_is_the_same_problems(ref):
    hp = W.profit(2048)
    sub = bench.strided(hp, 16, 128)
    assert sub.count == 128 and np.array_equal(sub.params, hp.params[::16])
    assert np.array_equal(ref.solve(sub, 2).num_iters, ref.solve(hp, 2).num_iters[::16])
    hl = W.lmi(32, general=True)
    sub = bench.strided(hl, 4, 8)
    assert np.array_equal(sub.xc0, hl.xc0[::4]) and sub.mq0 is hl.mq0          # shared arrays stay shared
    full, part = ref.solve(hl, 2), ref.solve(sub, 2)
    assert np.array_equal(part.num_iters, full.num_iters[::4]) and np.array_equal(part.exit_status, full.exit_status[::4])
    t = ref.lowpass_table(32)
    hw = W.lowpass(24, t, max_iters=200)
    sub = bench.strided(hw, 8, 3)
    assert np.array_equal(np.ravel(sub.gamma0), hw.gamma0[::8]) and np.array_equal(ref.solve(sub, 2).num_iters, ref.solve(hw, 2).num_iters[::8])


def test_iterations_done_counts_the_terminating_update():
    ni = np.array([25, 2000, 0], dtype=np.int64)
    es = np.array([4, 5, 1], dtype=np.int32)          # Tol, MaxIt, NoSoln
    assert bench.iterations_done(ni, es) == 25 + 2000 + 0 + 2   # MaxIt runs exactly max_iters updates, the others one more


def test_counters_file_is_present_and_sane():
    c = bench.load_counters()
    lp = c["lowpass32"]
    assert 300 < lp["shared_wavefronts_per_iteration"] < 600 and 400 < lp["fp64_warp_instr_per_iteration"] < 800
    assert 4.0e9 < c["sweep_n16384"]["dram_bytes_per_launch"] < 4.4e9
