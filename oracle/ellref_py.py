"""ctypes loader for the CPU checker (oracle/_build/libellref.so).

TEST INFRASTRUCTURE, NOT PRODUCT: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  Nothing under ellalgo_rs_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from ellalgo_rs_b200 import _abi  # noqa: E402  (shared plain-C records only)
from ellalgo_rs_b200.problem import BatchOutput, BatchProblem  # noqa: E402

_LIB_PATH = os.path.join(_HERE, "_build", "libellref.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "ellref.hpp", "ellref_oracles.hpp", "ellref_linalg.hpp")]
    srcs.append(os.path.join(_ROOT, "include", "ellcuda.h"))
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs if os.path.exists(s))
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


_pd = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.ellref_batch_solve.argtypes = [C.POINTER(_abi.BatchDesc), C.POINTER(_abi.BatchResult), C.c_int]
        L.ellref_batch_solve.restype = C.c_int
        L.ellref_space_new.argtypes = [C.c_int, C.c_int64, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ellref_space_new.restype = C.c_void_p
        L.ellref_space_free.argtypes = [C.c_void_p]
        L.ellref_space_clone.argtypes = [C.c_void_p]
        L.ellref_space_clone.restype = C.c_void_p
        L.ellref_space_update.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_double]
        L.ellref_space_update.restype = C.c_int
        L.ellref_space_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, _pd, _pd]
        L.ellref_space_set_xc.argtypes = [C.c_void_p, C.c_void_p]
        L.ellref_ell_update_inplace.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, _pd, _pd, C.c_int, C.c_int, C.c_int,
                                                C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_void_p]
        L.ellref_ell_update_inplace.restype = C.c_int
        L.ellref_calc.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double,
                                  C.c_double, C.c_void_p]
        L.ellref_calc.restype = C.c_int
        L.ellref_ldlt.argtypes = [C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, _pd]
        L.ellref_ldlt.restype = C.c_int
        L.ellref_lmi_assess.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _pd]
        L.ellref_lmi_assess.restype = C.c_int
        L.ellref_lmi_assess_variant.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _pd]
        L.ellref_lmi_assess_variant.restype = C.c_int
        L.ellref_dot.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.ellref_dot.restype = C.c_double
        L.ellref_dot_mv.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        L.ellref_dot_mv.restype = None
        L.ellref_linspace.argtypes = [C.c_double, C.c_double, C.c_int64, C.c_void_p]
        L.ellref_linspace.restype = None
        L.ellref_profit_assess.argtypes = [C.c_void_p, C.c_void_p, _pd, C.c_void_p, _pd]
        L.ellref_profit_assess.restype = C.c_int
        L.ellref_lowpass_table.argtypes = [C.c_int64, C.c_void_p]
        L.ellref_lowpass_band_edge.argtypes = [C.c_int64, C.c_double]
        L.ellref_lowpass_band_edge.restype = C.c_int
        L.ellref_lowpass_case_params.argtypes = [C.c_void_p]
        L.ellref_bsearch_example3.argtypes = [C.c_double, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int64,
                                              C.c_double, C.c_int64, C.POINTER(C.c_int)]
        L.ellref_bsearch_example3.restype = C.c_int64
        L.ellref_bsearch_bs2.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int64, C.POINTER(C.c_int)]
        L.ellref_bsearch_bs2.restype = C.c_int64
        L.ellref_hardware_threads.restype = C.c_int
        L.ellref_libm.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
        L.ellref_power_iteration.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, _pd]
        L.ellref_power_iteration.restype = C.c_int64
        L.ellref_conjugate_gradient.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int64]
        L.ellref_conjugate_gradient.restype = C.c_int64
        _lib = L
    return _lib


def solve(prob: BatchProblem, nthreads: int = 0) -> BatchOutput:
    """cutting_plane_* over every problem of `prob` on the host CPU; returns (output, threads used)."""
    assert not prob.on_device, "the CPU checker takes host arrays"
    out = BatchOutput.empty(prob.count, prob.n)
    d = prob.to_desc()
    r = out.to_result()
    rc = lib().ellref_batch_solve(C.byref(d), C.byref(r), nthreads)
    if rc < 0:
        raise RuntimeError(f"ellref_batch_solve failed: {rc}")
    out.threads = rc
    return out


def hardware_threads() -> int:
    return lib().ellref_hardware_threads()


class Space:
    """One reference search space (Ell or EllStable) stepped call by call."""

    def __init__(self, kind, n, kappa=1.0, mq=None, diag=None, xc=None, no_defer_trick=False, use_parallel_cut=True, _h=None):
        self.n, self.kind = n, kind
        if _h is not None:
            self._h = _h
            return
        xc = np.zeros(n) if xc is None else np.ascontiguousarray(xc, dtype=np.float64)
        mq = None if mq is None else np.ascontiguousarray(mq, dtype=np.float64)
        diag = None if diag is None else np.ascontiguousarray(diag, dtype=np.float64)
        self._h = lib().ellref_space_new(kind, n, kappa, None if mq is None else mq.ctypes.data,
                                         None if diag is None else diag.ctypes.data, xc.ctypes.data,
                                         int(no_defer_trick), int(use_parallel_cut))

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:     # module globals are already None at interpreter shutdown
            try:
                lib().ellref_space_free(self._h)
            except Exception:
                pass
            self._h = None

    def clone(self):
        return Space(self.kind, self.n, _h=lib().ellref_space_clone(self._h))

    def update(self, which, grad, beta0, beta1=None, single=True):
        g = np.ascontiguousarray(grad, dtype=np.float64)
        has = 1 if beta1 is not None else (-1 if single else 0)
        return lib().ellref_space_update(self._h, which, g.ctypes.data, float(beta0), has, 0.0 if beta1 is None else float(beta1))

    def state(self):
        xc, mq = np.zeros(self.n), np.zeros((self.n, self.n))
        k, t = C.c_double(), C.c_double()
        lib().ellref_space_get(self._h, xc.ctypes.data, mq.ctypes.data, C.byref(k), C.byref(t))
        return dict(xc=xc, mq=mq, kappa=k.value, tsq=t.value)

    def set_xc(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        lib().ellref_space_set_xc(self._h, x.ctypes.data)


def calc(n, fn, b0=0.0, b1=None, tsq=0.0, a3=0.0, a4=0.0, use_parallel_cut=True):
    out = np.zeros(3)
    st = lib().ellref_calc(n, int(use_parallel_cut), fn, b0, 1 if b1 is not None else 0, 0.0 if b1 is None else b1, tsq, a3, a4,
                           out.ctypes.data)
    return st, tuple(out)


def libm(which: int, x: np.ndarray) -> np.ndarray:
    """glibc exp (0) / log (1) elementwise — numpy's own SIMD exp/log are NOT libm's"""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    lib().ellref_libm(which, x.ctypes.data, y.ctypes.data, x.size)
    return y


def lowpass_table(n: int) -> np.ndarray:
    t = np.zeros(15 * n * n)
    lib().ellref_lowpass_table(n, t.ctypes.data)
    return t


def lowpass_case_params() -> np.ndarray:
    p = np.zeros(5)
    lib().ellref_lowpass_case_params(p.ctypes.data)
    return p


def power_iteration(variant: int, a: np.ndarray, x0: np.ndarray, max_iters: int, tol: float):
    """src/power_iteration.rs variants 1..4 → (eigenvalue, niter, x)"""
    a = np.ascontiguousarray(a, dtype=np.float64)
    x = np.array(x0, dtype=np.float64)
    eig = C.c_double()
    it = lib().ellref_power_iteration(variant, x.size, a.ctypes.data, x.ctypes.data, max_iters, tol, C.byref(eig))
    return eig.value, it, x


def conjugate_gradient(a: np.ndarray, b: np.ndarray, x0=None, tol: float = 1e-5, max_iter: int = 1000):
    """src/conjugate_gradient.rs → (x, iterations) ; iterations < 0 = the reference's Err cases"""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros(b.size) if x0 is None else np.array(x0, dtype=np.float64)
    it = lib().ellref_conjugate_gradient(b.size, a.ctypes.data, b.ctypes.data, x.ctypes.data, tol, max_iter)
    return x, it
