"""Host mirror of the reference's GEMV-sharing numerics: `power_iteration` (four variants,
src/power_iteration.rs:53-173) and `conjugate_gradient` (src/conjugate_gradient.rs:21-71), written against any matrix
object with `dot_mv(x)` — in particular ell.Ell, whose dot_mv is the same TMA-staged streaming pass Ell::update_core
uses.  Two ways to run them:

  * host vectors (numpy): every GEMV copies x in and y out; sums are sequential like the reference's (numpy's cumsum adds
    left to right);
  * `device=True` (or CUDA tensors passed in): the vectors stay on the GPU for the whole iteration — elementwise work as
    separately rounded torch kernels (no fused multiply-add: `a + s * b` is a multiply kernel and an add kernel), every sum
    by `ellcuda_seq_sum` in the reference's left-to-right order — and only the scalars the control flow needs come back.
Both give the checker's results bit for bit."""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


@dataclass
class Options:  # src/power_iteration.rs:16-19
    max_iters: int = 2000
    tolerance: float = 1e-9


_SUMMER = None    # the Ell whose device / stream run the ordered sums of device vectors (set by _prepare)


def _is_dev(v) -> bool:
    return type(v).__module__.startswith("torch")


def _seq_sum(v) -> float:
    if _is_dev(v):
        return _SUMMER.seq_sum(v.contiguous())
    return float(np.cumsum(v)[-1]) if v.size else 0.0


def _dot(a, b) -> float:  # Arr::dot, src/arr.rs:443-451
    return _seq_sum(a * b)


def _norm_l1(x) -> float:  # src/power_iteration.rs:22-25
    return _seq_sum(x.abs() if _is_dev(x) else np.abs(x))


def _prepare(a, x, device: bool):
    """the vector as the iteration will carry it: a float64 CUDA tensor on the matrix' device (device path) or a numpy array"""
    global _SUMMER
    if device or _is_dev(x):
        import torch
        _SUMMER = a
        return x.to(torch.float64) if _is_dev(x) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(f"cuda:{a.device}")
    return np.asarray(x, dtype=np.float64)


def _finish(x):
    """results are returned as numpy arrays either way"""
    return x.cpu().numpy() if _is_dev(x) else x


def _div(x, s):  # Arr /= f64 multiplies by the reciprocal, src/arr.rs:241-249
    return x * (1.0 / s)


def power_iteration(a, x, options: Options, device: bool = False):  # :53-65 → (eigenvalue, niter, x)
    x = _prepare(a, x, device)
    x = _div(x, math.sqrt(_dot(x, x)))
    for niter in range(options.max_iters):
        x1 = x
        x = a.dot_mv(x1)
        x = _div(x, math.sqrt(_dot(x, x)))
        if _norm_l1(x - x1) <= options.tolerance or _norm_l1(x + x1) <= options.tolerance:
            return _dot(x, a.dot_mv(x)), niter, _finish(x)
    return _dot(x, a.dot_mv(x)), options.max_iters, _finish(x)


def power_iteration4(a, x, options: Options, device: bool = False):  # :81-96
    x = _prepare(a, x, device)
    x = _div(x, _norm_l1(x))
    for niter in range(options.max_iters):
        x1 = x
        x = a.dot_mv(x1)
        x = _div(x, _norm_l1(x))
        if _norm_l1(x - x1) <= options.tolerance or _norm_l1(x + x1) <= options.tolerance:
            x = _div(x, math.sqrt(_dot(x, x)))
            return _dot(x, a.dot_mv(x)), niter, _finish(x)
    x = _div(x, math.sqrt(_dot(x, x)))
    return _dot(x, a.dot_mv(x)), options.max_iters, _finish(x)


def power_iteration2(a, x, options: Options, device: bool = False):  # :112-127
    x = _prepare(a, x, device)
    x = _div(x, math.sqrt(_dot(x, x)))
    new_vec = a.dot_mv(x)
    eigenval = _dot(x, new_vec)
    for niter in range(options.max_iters):
        prev = eigenval
        x = _div(new_vec, math.sqrt(_dot(new_vec, new_vec)))
        new_vec = a.dot_mv(x)
        eigenval = _dot(x, new_vec)
        if abs(prev - eigenval) <= options.tolerance:
            return eigenval, niter, _finish(x)
    return eigenval, options.max_iters, _finish(x)


def power_iteration3(a, x, options: Options, device: bool = False):  # :144-173
    x = _prepare(a, x, device)
    new_vec = a.dot_mv(x)
    dot = _dot(x, x)
    eigenval = _dot(x, new_vec) / dot
    for niter in range(options.max_iters):
        prev = eigenval
        x = new_vec
        dot = _dot(x, x)
        if dot >= 1e150:
            x = _div(x, math.sqrt(_dot(x, x)))
            new_vec = a.dot_mv(x)
            eigenval = _dot(x, new_vec)
            if abs(prev - eigenval) <= options.tolerance:
                return eigenval, niter, _finish(x)
        else:
            new_vec = a.dot_mv(x)
            eigenval = _dot(x, new_vec) / dot
            if abs(prev - eigenval) <= options.tolerance:
                return eigenval, niter, _finish(_div(x, math.sqrt(_dot(x, x))))
    return eigenval, options.max_iters, _finish(x)


class ConjGradError(RuntimeError):
    pass


def conjugate_gradient(a, b, x0=None, tol: float = 1e-5, max_iter: int = 1000, device: bool = False):  # src/conjugate_gradient.rs:21-71
    """returns x; raises ConjGradError where the reference returns Err(...)"""
    b = _prepare(a, b, device)
    n = b.numel() if _is_dev(b) else b.size
    x = _prepare(a, np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64), device or _is_dev(b))
    residual = b - a.dot_mv(x)
    direction = residual.clone() if _is_dev(residual) else residual.copy()
    rns = _dot(residual, residual)
    for _ in range(max_iter):
        a_dir = a.dot_mv(direction)
        dad = _dot(direction, a_dir)
        if dad == 0.0:
            raise ConjGradError("Conj Grad did not converge")
        step = rns / dad
        x = x + step * direction
        residual = residual - step * a_dir
        rns_new = _dot(residual, residual)
        if math.sqrt(rns_new) < tol:
            return _finish(x)
        ratio = rns_new / rns
        direction = residual + ratio * direction
        rns = rns_new
    raise ConjGradError("Conj Grad did not converge after max iterations")
