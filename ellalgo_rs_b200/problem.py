"""Host-side description of a batch of independent cutting-plane problems.

`BatchProblem` is the Python picture of `ell_batch_desc` (include/ellcuda.h): per problem i it stands for

    space = Ell::new_with_matrix(kappa0[i], mq0 | diag(diag0[i]) | I, xc0[i])     src/ell.rs:31-78
    omega = <Oracle>::new(params[i], table)
    cutting_plane_{optim,feas,optim_q}(omega, space, gamma0[i], Options{max_iters, tolerance})

Arrays may be numpy (host) or torch CUDA tensors (device); a 1-D / 0-D array where a per-instance one is
expected means "shared by every instance" (stride 0).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any, Optional, Sequence

import numpy as np

from . import _abi


def _is_torch(a) -> bool:
    return type(a).__module__.startswith("torch")


def _ptr(a) -> int:
    if a is None:
        return 0
    if _is_torch(a):
        return a.data_ptr()
    return a.ctypes.data


def _prep(a, dtype=np.float64):
    """contiguous float64 (numpy) / leave torch tensors alone (must already be contiguous f64)."""
    if a is None:
        return None
    if _is_torch(a):
        import torch
        assert a.dtype == torch.float64 and a.is_contiguous(), "device arrays must be contiguous float64"
        return a
    return np.ascontiguousarray(np.asarray(a, dtype=dtype))


@dataclass
class BatchProblem:
    oracle_id: int
    n: int
    count: int
    mode: int = _abi.MODE_OPTIM
    kind: int = _abi.KIND_ELL
    max_iters: int = 2000            # Options::default, src/cutting_plane.rs:97-99
    tolerance: float = 1e-20
    no_defer_trick: bool = False
    use_parallel_cut: bool = True
    kappa0: Any = 1.0                # scalar, or (count,)
    xc0: Any = None                  # (n,) shared or (count, n); None ⇒ zeros
    diag0: Any = None                # (n,) / (count, n); None ⇒ identity
    mq0: Any = None                  # (n, n) / (count, n, n)
    gamma0: Any = 0.0                # scalar or (count,)
    params: Any = None               # (w,) shared or (count, w)
    table: Any = None                # 1-D shared table
    aux: Sequence[int] = field(default_factory=list)
    device: int = 0

    def __post_init__(self):
        if self.xc0 is None:
            self.xc0 = np.zeros(self.n)
        for name in ("kappa0", "xc0", "diag0", "mq0", "gamma0", "params", "table"):
            v = getattr(self, name)
            if v is not None and not _is_torch(v):
                setattr(self, name, _prep(np.atleast_1d(np.asarray(v, dtype=np.float64))))
        self._check()

    # -- shape rules ------------------------------------------------------------------------
    def _stride(self, a, width: int) -> int:
        if a is None:
            return 0
        size = a.numel() if _is_torch(a) else a.size
        if size == width:
            return 0
        if size == width * self.count:
            return width
        raise ValueError(f"array of {size} values does not match width {width} x count {self.count}")

    def _check(self):
        n = self.n
        if n <= 0 or self.count < 0:
            raise ValueError("n must be positive and count non-negative")
        self._stride(self.kappa0, 1)
        self._stride(self.xc0, n)
        self._stride(self.diag0, n)
        self._stride(self.mq0, n * n)
        self._stride(self.gamma0, 1)
        if self.params is not None:
            self.params_width  # noqa: B018  (validates)

    @property
    def params_width(self) -> int:
        if self.params is None:
            return 0
        p = self.params
        if p.ndim == 1:
            return p.shape[0]
        if p.ndim == 2 and p.shape[0] in (1, self.count):
            return p.shape[1]
        raise ValueError("params must be (w,) or (count, w)")

    @property
    def on_device(self) -> bool:
        return any(_is_torch(getattr(self, k)) for k in ("kappa0", "xc0", "diag0", "mq0", "gamma0", "params", "table")
                   if getattr(self, k) is not None)

    def to_desc(self, stream: int = 0) -> _abi.BatchDesc:
        d = _abi.BatchDesc()
        d.kind, d.oracle_id, d.mode, d.n = self.kind, self.oracle_id, self.mode, self.n
        d.count, d.max_iters, d.tolerance = self.count, self.max_iters, self.tolerance
        d.no_defer_trick, d.use_parallel_cut = int(self.no_defer_trick), int(self.use_parallel_cut)
        n = self.n
        d.kappa0, d.kappa0_stride = _ptr(self.kappa0), self._stride(self.kappa0, 1)
        d.xc0, d.xc0_stride = _ptr(self.xc0), self._stride(self.xc0, n)
        d.diag0, d.diag0_stride = _ptr(self.diag0), self._stride(self.diag0, n)
        d.mq0, d.mq0_stride = _ptr(self.mq0), self._stride(self.mq0, n * n)
        d.gamma0, d.gamma0_stride = _ptr(self.gamma0), self._stride(self.gamma0, 1)
        w = self.params_width
        d.params, d.params_width = _ptr(self.params), w
        d.params_stride = 0 if (self.params is None or self.params.ndim == 1 or self.params.shape[0] == 1) else w
        d.table = _ptr(self.table)
        d.table_len = 0 if self.table is None else (self.table.numel() if _is_torch(self.table) else self.table.size)
        for k, v in enumerate(self.aux):
            d.aux[k] = int(v)
        d.device = self.device
        d.on_device = int(self.on_device)
        d.stream = stream
        return d

    def slice(self, lo: int, hi: int) -> "BatchProblem":
        """problems [lo, hi) — used to shard a batch across ranks (no data-path collective)."""
        def cut(a, width):
            if a is None:
                return None
            return a if self._stride(a, width) == 0 else a.reshape(self.count, -1)[lo:hi]
        n = self.n
        return BatchProblem(
            oracle_id=self.oracle_id, n=n, count=hi - lo, mode=self.mode, kind=self.kind, max_iters=self.max_iters,
            tolerance=self.tolerance, no_defer_trick=self.no_defer_trick, use_parallel_cut=self.use_parallel_cut,
            kappa0=cut(self.kappa0, 1), xc0=cut(self.xc0, n), diag0=cut(self.diag0, n), mq0=cut(self.mq0, n * n),
            gamma0=cut(self.gamma0, 1),
            params=None if self.params is None else (self.params if self.params.ndim == 1 else self.params[lo:hi]),
            table=self.table, aux=list(self.aux), device=self.device)


@dataclass
class BatchOutput:
    """Python picture of `ell_batch_result` (host numpy arrays)."""
    xbest: np.ndarray        # (count, n)
    has_xbest: np.ndarray    # (count,) uint8
    num_iters: np.ndarray    # (count,) int64
    exit_status: np.ndarray  # (count,) int32
    gamma: np.ndarray        # (count,)
    tsq: np.ndarray          # (count,)
    xc: np.ndarray           # (count, n)
    kappa: np.ndarray        # (count,)

    @staticmethod
    def empty(count: int, n: int) -> "BatchOutput":
        return BatchOutput(
            xbest=np.zeros((count, n)), has_xbest=np.zeros(count, np.uint8), num_iters=np.zeros(count, np.int64),
            exit_status=np.zeros(count, np.int32), gamma=np.zeros(count), tsq=np.zeros(count),
            xc=np.zeros((count, n)), kappa=np.zeros(count))

    def to_result(self) -> _abi.BatchResult:
        r = _abi.BatchResult()
        for name in ("xbest", "has_xbest", "num_iters", "exit_status", "gamma", "tsq", "xc", "kappa"):
            setattr(r, name, getattr(self, name).ctypes.data)
        return r

    def concat(self, other: "BatchOutput") -> "BatchOutput":
        return BatchOutput(*[np.concatenate([getattr(self, f), getattr(other, f)]) for f in
                             ("xbest", "has_xbest", "num_iters", "exit_status", "gamma", "tsq", "xc", "kappa")])
