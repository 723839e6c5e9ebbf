"""Run by tests/test_ell64_gpu.py::test_tmem_kernel_bowl_and_lowpass_when_selected in a child process with
ELLCUDA_CTA_MIN_N=65 (the library reads its tuning switches once per process): the tensor-memory kernel's bowl and lowpass
paths for 32 < n <= 64, which the default routing gives to the CTA-per-problem kernel, against the CPU checker."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402

import ellref_py as ref  # noqa: E402
from ellalgo_rs_b200 import batch  # noqa: E402
from ellalgo_rs_b200 import workloads as W  # noqa: E402
from parity import assert_parity  # noqa: E402

assert os.environ.get("ELLCUDA_CTA_MIN_N") == "65"
ref.build()
for n, count, max_iters in [(33, 40, 500), (48, 65, 700), (64, 130, 900)]:
    prob = W.bowl(count, n=n, max_iters=max_iters)
    assert_parity(batch.solve_batch(prob), ref.solve(prob), exact=True, what=f"TMEM bowl n={n}")
n = 50
rng = np.random.default_rng(5)
prob = W.bowl(19, n=n, max_iters=200)
prob.kappa0 = np.ones(1)
prob.mq0 = np.ascontiguousarray(100.0 * (np.eye(n) + rng.uniform(-1, 1, (19, n, n)) / n))   # non-symmetric start
for nd in (False, True):
    prob.no_defer_trick = nd
    assert_parity(batch.solve_batch(prob), ref.solve(prob), exact=True, what=f"TMEM bowl n={n} matrix start no_defer={nd}")
for n, count, max_iters in [(64, 7, 1200), (52, 5, 800)]:
    prob = W.lowpass(count, batch.lowpass_table(n), n=n, max_iters=max_iters)
    assert_parity(batch.solve_batch(prob), ref.solve(prob), exact=True, what=f"TMEM lowpass n={n}")
print("tmem kernel ok")
