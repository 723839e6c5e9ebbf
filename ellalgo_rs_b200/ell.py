"""`Ell` — host mirror of the reference's search space (src/ell.rs) whose matrix lives in B200 HBM and whose
update_core (src/ell.rs:97-137) runs as the fused TMA sweep of csrc/stream_update.cu.  Same constructors and
SearchSpace methods as the Rust type; every call goes through the C ABI (include/ellcuda.h)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from ._lib import check, lib
from .cutting_plane import CutStatus, ParallelCut, SingleCut


def _cut_args(cut):
    if isinstance(cut, SingleCut):
        return float(cut.beta), -1, 0.0
    if isinstance(cut, ParallelCut):
        if cut.beta1 is None:
            return float(cut.beta0), 0, 0.0
        return float(cut.beta0), 1, float(cut.beta1)
    raise TypeError("cut must be SingleCut or ParallelCut")


class _DevVec:
    """zero-copy view of a device buffer owned by the handle (for torch.as_tensor / all_gather)"""

    def __init__(self, ptr: int, n: int, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}


class Ell:
    KIND = _abi.KIND_ELL

    def __init__(self, kappa, xc, mq=None, diag=None, *, device=0, no_defer_trick=False, use_parallel_cut=True,
                 rows=None, reduce_mode=0, stream=0, devices=None, storage=0, shard=None, _handle=None):
        """devices=[d0, d1, …]: rows sharded over these devices of this process (ellcuda_create_sharded); the object is
        then used exactly like a single-device one."""
        self._keep = []
        if _handle is not None:
            self._h = _handle
            return
        xc = np.ascontiguousarray(xc, dtype=np.float64)
        n = xc.shape[0]
        d = _abi.CreateDesc()
        d.kind, d.device, d.n, d.kappa = self.KIND, device, n, float(kappa)
        if mq is not None:
            mq = np.ascontiguousarray(mq, dtype=np.float64)
            if mq.shape != (n, n):
                raise ValueError("mq must be n x n")  # the reference asserts (src/arr.rs:427-429)
            d.mq = mq.ctypes.data
        if diag is not None:
            diag = np.ascontiguousarray(diag, dtype=np.float64)
            if diag.shape != (n,):
                raise ValueError("diag must have n entries")
            d.diag = diag.ctypes.data
        d.xc = xc.ctypes.data
        d.no_defer_trick, d.use_parallel_cut = int(no_defer_trick), int(use_parallel_cut)
        d.row_begin, d.row_end = (0, n) if rows is None else rows
        d.reduce_mode = reduce_mode
        d.storage = storage          # 0 = full matrix (bit-identical), 1 = lower block-triangle only (half the traffic, 1e-9)
        d.stream = stream
        if shard is not None:        # symmetric storage sharded over ranks: (rank, count)
            d.shard_rank, d.shard_count = shard
        h = C.c_void_p()
        if devices is not None:
            devs = (C.c_int32 * len(devices))(*devices)
            d.row_begin, d.row_end = 0, n
            check(lib().ellcuda_create_sharded(C.byref(d), len(devices), devs, C.byref(h)))
            device = devices[0]
        else:
            check(lib().ellcuda_create(C.byref(d), C.byref(h)))
        self._h = h
        self.n = n
        self.rows = (d.row_begin, d.row_end)
        self.device = device
        self._no_defer_trick = bool(no_defer_trick)
        self._use_parallel_cut = bool(use_parallel_cut)

    # ---- the reference's public knobs, writable after construction (src/ell.rs:10, 342-344; src/ell_calc.rs:630)
    @property
    def no_defer_trick(self) -> bool:
        return self._no_defer_trick

    @no_defer_trick.setter
    def no_defer_trick(self, on: bool):
        check(lib().ellcuda_set_no_defer_trick(self._h, int(bool(on))))
        self._no_defer_trick = bool(on)

    @property
    def use_parallel_cut(self) -> bool:
        return self._use_parallel_cut

    @use_parallel_cut.setter
    def use_parallel_cut(self, on: bool):
        check(lib().ellcuda_set_use_parallel_cut(self._h, int(bool(on))))
        self._use_parallel_cut = bool(on)

    @property
    def group_size(self) -> int:
        return int(lib().ellcuda_group_size(self._h))

    # ---- constructors, src/ell.rs:31-78
    @classmethod
    def new_with_matrix(cls, kappa, mq, xc, **kw):
        return cls(kappa, xc, mq=mq, **kw)

    @classmethod
    def new(cls, val, xc, **kw):
        return cls(1.0, xc, diag=val, **kw)

    @classmethod
    def new_with_scalar(cls, val, xc, **kw):
        return cls(val, xc, **kw)

    @classmethod
    def from_covariance(cls, cov, xc, **kw):
        return cls(1.0, xc, mq=cov, **kw)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().ellcuda_destroy(h)
            except Exception:
                pass
            self._h = None

    def clone(self):  # #[derive(Clone)]
        h = C.c_void_p()
        check(lib().ellcuda_clone(self._h, C.byref(h)))
        e = type(self)(0, None, _handle=h)
        e.n, e.rows, e.device = self.n, self.rows, self.device
        e._no_defer_trick, e._use_parallel_cut = self._no_defer_trick, self._use_parallel_cut
        return e

    # ---- SearchSpace, src/ell.rs:140-180
    def xc(self) -> np.ndarray:
        out = np.empty(self.n)
        check(lib().ellcuda_xc(self._h, out.ctypes.data))
        return out

    def set_xc(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.shape != (self.n,):
            raise ValueError("xc must have n entries")
        check(lib().ellcuda_set_xc(self._h, x.ctypes.data))

    def tsq(self) -> float:
        v = C.c_double()
        check(lib().ellcuda_tsq(self._h, C.byref(v)))
        return v.value

    @property
    def kappa(self) -> float:
        v = C.c_double()
        check(lib().ellcuda_kappa(self._h, C.byref(v)))
        return v.value

    @property
    def mq(self) -> np.ndarray:
        """the rows this handle holds, (row_end-row_begin, n)"""
        out = np.empty((self.rows[1] - self.rows[0], self.n))
        check(lib().ellcuda_get_mq(self._h, out.ctypes.data))
        return out

    def _update(self, which, cut):
        grad, beta = cut
        grad = np.ascontiguousarray(grad, dtype=np.float64)
        if grad.shape != (self.n,):
            raise ValueError("grad must have n entries")  # reference: assert_eq!(self.cols, x.len())
        b0, has, b1 = _cut_args(beta)
        st = C.c_int32()
        check(lib().ellcuda_update(self._h, which, grad.ctypes.data, b0, has, b1, C.byref(st)))
        return CutStatus(st.value)

    def update_bias_cut(self, cut):
        return self._update(_abi.UPDATE_BIAS, cut)

    def update_central_cut(self, cut):
        return self._update(_abi.UPDATE_CENTRAL, cut)

    def update_q(self, cut):
        return self._update(_abi.UPDATE_Q, cut)

    # ---- Arr::dot_mv on the device-resident matrix (src/arr.rs:426-442)
    def dot_mv_device(self, x_dev_ptr: int, y_dev_ptr: int):
        check(lib().ellcuda_dot_mv(self._h, x_dev_ptr, y_dev_ptr))

    def seq_sum(self, v_dev) -> float:
        """sum of a CUDA tensor strictly left to right (the reference's Arr::dot / sum order) on this handle's device"""
        import torch
        torch.cuda.current_stream(self.device).synchronize()          # v_dev was produced on torch's stream
        out = C.c_double()
        check(lib().ellcuda_seq_sum(self._h, v_dev.data_ptr(), v_dev.numel(), C.byref(out)))
        return out.value

    def dot_mv(self, x):
        """y = mq·x; sums in the reference's order.  A host vector goes H2D / streaming GEMV / D2H; a CUDA tensor on this
        handle's device stays on the device and a CUDA tensor comes back (linalg.py keeps whole iterations on the GPU that way)"""
        import torch
        if isinstance(x, torch.Tensor):
            if x.dtype != torch.float64 or x.numel() != self.n or x.device.index != self.device:
                raise ValueError("x must be a float64 CUDA tensor of n entries on the handle's device")
            x = x.contiguous()
            y = torch.empty(self.n if self.group_size > 1 else self.rows[1] - self.rows[0], dtype=torch.float64, device=x.device)
            torch.cuda.current_stream(self.device).synchronize()      # x may have been produced on torch's stream
            self.dot_mv_device(x.data_ptr(), y.data_ptr())
            self.sync()
            return y
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.shape != (self.n,):
            raise ValueError("x must have n entries")  # assert_eq!(self.cols, x.len())
        dev = f"cuda:{self.device}"          # the handle's device, not torch's current one
        xd = torch.from_numpy(x).to(dev)
        yd = torch.empty(self.rows[1] - self.rows[0], dtype=torch.float64, device=dev)
        torch.cuda.synchronize(self.device)
        self.dot_mv_device(xd.data_ptr(), yd.data_ptr())
        self.sync()
        return yd.cpu().numpy()

    # ---- split form for device-resident gradients / sharded slabs / CUDA graphs
    def update_begin(self, which, grad_dev_ptr: int, beta):
        b0, has, b1 = _cut_args(beta)
        check(lib().ellcuda_update_begin(self._h, which, grad_dev_ptr, b0, has, b1))

    def update_finish(self):
        check(lib().ellcuda_update_finish(self._h))

    def last_status(self) -> CutStatus:
        st = C.c_int32()
        check(lib().ellcuda_last_status(self._h, C.byref(st)))
        return CutStatus(st.value)

    def grad_t_view(self) -> _DevVec:
        return _DevVec(lib().ellcuda_grad_t_ptr(self._h), self.n, self)

    # ---- peer exchange over NVLink (include/ellcuda.h: ellcuda_exchange_*)
    def exchange_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        check(lib().ellcuda_exchange_export(self._h, buf))
        return buf.raw

    def exchange_attach(self, handles: list, rank: int):
        """handles[r] = exchange_export() of rank r (any order of arrival; own entry is ignored)"""
        blob = b"".join(handles)
        check(lib().ellcuda_exchange_attach(self._h, len(handles), rank, blob))

    def exchange_block_ptr(self) -> int:
        return lib().ellcuda_exchange_block_ptr(self._h)

    def exchange_attach_ptrs(self, blocks: list, rank: int):
        arr = (C.c_void_p * len(blocks))(*blocks)
        check(lib().ellcuda_exchange_attach_ptrs(self._h, len(blocks), rank, arr))

    def set_stream(self, stream: int):
        check(lib().ellcuda_set_stream(self._h, stream))

    def sync(self):
        check(lib().ellcuda_sync(self._h))

    def set_timing(self, on: bool):
        check(lib().ellcuda_set_timing(self._h, int(on)))

    def last_sweep_ms(self) -> float:
        return float(lib().ellcuda_last_sweep_ms(self._h))

    @property
    def launches(self) -> int:
        return int(lib().ellcuda_handle_launch_count(self._h))


class EllStable(Ell):
    """Host mirror of the reference's `EllStable` (src/ell_stable.rs) as a single host-driven search space, n <= 2048:
    `new`, `new_with_scalar`, `new_with_matrix` (:18-35) and the SearchSpace methods (:128-166); update_core (:52-125)
    runs as one single-warp kernel per call.  `mq` returns the factor (diagonal + strict upper triangle)."""
    KIND = _abi.KIND_ELLSTABLE

    from_covariance = None   # the reference's EllStable has no such constructor
