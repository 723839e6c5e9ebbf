"""Parity gate shared by the GPU tests: per instance identical num_iters, exit status and
has_xbest; xbest and final gamma within REL_TOL = 1e-9 relative (north_star's f64 tolerance) — or bit-identical
where the path has no libm call."""
import numpy as np

REL_TOL = 1e-9


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    both_nan = np.isnan(a) & np.isnan(b)
    with np.errstate(invalid="ignore", divide="ignore"):
        e = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    e = np.where(both_inf | both_nan | (a == b), 0.0, e)
    return e


def assert_parity(gpu, ref, exact: bool, what: str = "", tol: float = REL_TOL):
    assert np.array_equal(gpu.num_iters, ref.num_iters), (
        f"{what}: num_iters differ at {np.flatnonzero(gpu.num_iters != ref.num_iters)[:8]}: "
        f"gpu {gpu.num_iters[gpu.num_iters != ref.num_iters][:8]} ref {ref.num_iters[gpu.num_iters != ref.num_iters][:8]}")
    assert np.array_equal(gpu.exit_status, ref.exit_status), f"{what}: exit status differs"
    assert np.array_equal(gpu.has_xbest, ref.has_xbest), f"{what}: has_xbest differs"
    has = ref.has_xbest.astype(bool)
    if exact:
        assert np.array_equal(gpu.xbest[has], ref.xbest[has]), f"{what}: xbest not bit-identical"
        for f in ("gamma", "tsq", "kappa", "xc"):
            a, b = getattr(gpu, f), getattr(ref, f)
            assert np.array_equal(a, b, equal_nan=True), f"{what}: {f} not bit-identical (max rel {rel_err(a, b).max():.3g})"
    else:
        assert rel_err(gpu.xbest[has], ref.xbest[has]).max(initial=0.0) <= tol, f"{what}: xbest beyond {tol}"
        assert rel_err(gpu.gamma, ref.gamma).max(initial=0.0) <= tol, f"{what}: gamma beyond {tol}"
