"""Committed golden vectors (tests/golden/golden_v1.npz, produced by tests/golden/make_golden.py): the CPU checker
must still reproduce them (CPU test) and the CUDA path must reproduce them bit for bit (GPU test)."""
import os

import numpy as np
import pytest

from ellalgo_rs_b200 import _abi as A
from golden_cases import CASES, CASES_V2, CASES_V3, run_stepwise, run_stepwise_nonsym

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))
GOLD2 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v2.npz"))
GOLD3 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v3.npz"))
FIELDS = ("xbest", "has_xbest", "num_iters", "exit_status", "gamma", "tsq", "xc", "kappa")


def _check(name, o, gold=None):
    gold = GOLD if gold is None else gold
    for f in FIELDS:
        assert np.array_equal(getattr(o, f), gold[f"{name}/{f}"], equal_nan=True), f"{name}/{f} differs from the golden vector"


def test_golden_has_the_pinned_counts():
    assert GOLD["example1/num_iters"][0] == 25 and GOLD["profit/num_iters"][0] == 83 and GOLD["profit_q/num_iters"][0] == 29
    assert GOLD["lowpass32/num_iters"][0] == 12481


@pytest.mark.parametrize("name", list(CASES))
def test_checker_reproduces_golden(ref, name):
    _check(name, ref.solve(CASES[name](ref)))


def test_checker_stepwise_golden(ref):
    st = run_stepwise(lambda **kw: ref.Space(A.KIND_ELL, **kw), lambda s, w, g, b0, b1: s.update(w, g, b0, b1, single=False),
                      lambda s: s.state())
    for k, v in st.items():
        assert np.array_equal(v, GOLD[f"stepwise/{k}"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_reproduces_golden(ref, name):
    from ellalgo_rs_b200 import batch
    _check(name, batch.solve_batch(CASES[name](ref)))


@pytest.mark.gpu
def test_cuda_stepwise_golden():
    from ellalgo_rs_b200.cutting_plane import ParallelCut
    from ellalgo_rs_b200.ell import Ell
    st = run_stepwise(lambda n, kappa, mq, xc: Ell(kappa, xc, mq=mq),
                      lambda s, w, g, b0, b1: s.update_bias_cut((g, ParallelCut(b0, b1))),
                      lambda s: dict(xc=s.xc(), mq=s.mq, kappa=s.kappa, tsq=s.tsq()))
    for k, v in st.items():
        assert np.array_equal(v, GOLD[f"stepwise/{k}"])


# ---------------------------------------------------------------- round 2 vectors (golden_v2.npz)
@pytest.mark.parametrize("name", list(CASES_V2))
def test_checker_reproduces_golden_v2(ref, name):
    _check(name, ref.solve(CASES_V2[name](ref)), GOLD2)


def test_checker_stepwise_nonsym_golden(ref):
    st = run_stepwise_nonsym(lambda **kw: ref.Space(A.KIND_ELL, **kw), lambda s, w, g, b0, b1: s.update(w, g, b0, b1, single=False),
                             lambda s: s.state())
    for k, v in st.items():
        assert np.array_equal(v, GOLD2[f"stepwise_nonsym/{k}"])
    assert not np.array_equal(GOLD2["stepwise_nonsym/mq_before_first_success"], GOLD2["stepwise_nonsym/mq_before_first_success"].T)
    assert np.array_equal(GOLD2["stepwise_nonsym/mq"], GOLD2["stepwise_nonsym/mq"].T)      # mirrored by the first successful cut


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES_V2))
def test_cuda_reproduces_golden_v2(ref, name):
    from ellalgo_rs_b200 import batch
    _check(name, batch.solve_batch(CASES_V2[name](ref)), GOLD2)


@pytest.mark.gpu
def test_cuda_stepwise_nonsym_golden():
    from ellalgo_rs_b200.cutting_plane import ParallelCut
    from ellalgo_rs_b200.ell import Ell
    st = run_stepwise_nonsym(lambda n, kappa, mq, xc: Ell(kappa, xc, mq=mq),
                             lambda s, w, g, b0, b1: s.update_bias_cut((g, ParallelCut(b0, b1))),
                             lambda s: dict(xc=s.xc(), mq=s.mq, kappa=s.kappa, tsq=s.tsq()))
    for k, v in st.items():
        assert np.array_equal(v, GOLD2[f"stepwise_nonsym/{k}"])


# ---------------------------------------------------------------- batched Ell above n = 64 (golden_v3.npz)
@pytest.mark.parametrize("name", list(CASES_V3))
def test_checker_reproduces_golden_v3(ref, name):
    _check(name, ref.solve(CASES_V3[name](ref)), GOLD3)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES_V3))
def test_cuda_reproduces_golden_v3(ref, name):
    from ellalgo_rs_b200 import batch
    _check(name, batch.solve_batch(CASES_V3[name](ref)), GOLD3)
