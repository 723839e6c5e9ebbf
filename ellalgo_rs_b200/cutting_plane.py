"""Host mirror of the reference's plugin surface (src/cutting_plane.rs), same names and argument meaning:

    SingleCut / ParallelCut     :9, :18          CutStatus :32-37        Options :50-100
    cutting_plane_feas :205-227   cutting_plane_optim :286-313   cutting_plane_optim_q :331-374
    BSearchAdaptor :376-419       bsearch :441-466

A search space is anything with xc() / tsq() / update_bias_cut / update_central_cut / update_q / set_xc
(the SearchSpace trait, :154-182) — here ell.Ell, whose updates run on the GPU.  An oracle is any object with
assess_feas(xc) / assess_optim(xc, gamma) / assess_optim_q(xc, gamma, retry) (:119-147); because Python has no
`&mut f64`, assess_optim returns ((grad, cut), shrunk, gamma) and assess_optim_q returns
((grad, cut), shrunk, x_q, more_alt, gamma).
"""
from __future__ import annotations

import copy
import enum
from dataclasses import dataclass
from typing import Optional


@dataclass(frozen=True)
class SingleCut:  # src/cutting_plane.rs:9
    beta: float


@dataclass(frozen=True)
class ParallelCut:  # src/cutting_plane.rs:18
    beta0: float
    beta1: Optional[float] = None


class CutStatus(enum.IntEnum):  # src/cutting_plane.rs:32-37
    Success = 0
    NoSoln = 1
    NoEffect = 2
    Unknown = 3


@dataclass
class Options:  # src/cutting_plane.rs:50-100
    max_iters: int = 2000
    tolerance: float = 1e-20
    verbose: bool = False


def cutting_plane_feas(omega, space, options: Options = Options()):  # :205-227
    for niter in range(options.max_iters):
        cut = omega.assess_feas(space.xc())
        if cut is None:
            return space.xc(), niter
        status = space.update_bias_cut(cut)
        if status != CutStatus.Success or space.tsq() < options.tolerance:
            return None, niter
    return None, options.max_iters


def cutting_plane_optim(omega, space, gamma: float, options: Options = Options()):  # :286-313
    """returns (x_best, num_iters, gamma) — gamma is the reference's in-out `&mut f64`"""
    x_best = None
    for niter in range(options.max_iters):
        cut, shrunk, gamma = omega.assess_optim(space.xc(), gamma)
        if shrunk:
            x_best = space.xc()
            status = space.update_central_cut(cut)
        else:
            status = space.update_bias_cut(cut)
        if status != CutStatus.Success or space.tsq() < options.tolerance:
            return x_best, niter, gamma
    return x_best, options.max_iters, gamma


def cutting_plane_optim_q(omega, space_q, gamma: float, options: Options = Options()):  # :331-374
    x_best = None
    retry = False
    for niter in range(options.max_iters):
        cut, shrunk, x_q, more_alt, gamma = omega.assess_optim_q(space_q.xc(), gamma, retry)
        if shrunk:
            x_best = x_q
            retry = False
        status = space_q.update_q(cut)
        if status == CutStatus.Success:
            retry = False
        elif status == CutStatus.NoSoln:
            return x_best, niter, gamma
        elif status == CutStatus.NoEffect:
            if not more_alt:
                return x_best, niter, gamma
            retry = True
        if space_q.tsq() < options.tolerance:
            return x_best, niter, gamma
    return x_best, options.max_iters, gamma


class BSearchAdaptor:  # :376-419
    def __init__(self, omega, space, options: Options = Options()):
        self.omega, self.space, self.options = omega, space, options

    def assess_bs(self, gamma: float) -> bool:
        space = self.space.clone() if hasattr(self.space, "clone") else copy.deepcopy(self.space)
        self.omega.update(gamma)
        x_feas, _ = cutting_plane_feas(self.omega, space, self.options)
        if x_feas is not None:
            self.space.set_xc(x_feas)
            return True
        return False


def bsearch(omega, intrvl, options: Options = Options()):  # :441-466
    lower, upper = intrvl
    assert lower <= upper
    u_orig = upper
    for niter in range(options.max_iters):
        tau = (upper - lower) / 2.0
        if tau < options.tolerance:
            return upper != u_orig, niter
        gamma = lower
        gamma += tau
        if omega.assess_bs(gamma):
            upper = gamma
        else:
            lower = gamma
    return upper != u_orig, options.max_iters
