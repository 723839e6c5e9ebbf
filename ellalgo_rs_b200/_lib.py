"""ctypes binding of csrc/libellcuda.so (include/ellcuda.h).  Loading works without a GPU (so the symbol table can
be checked on a CPU box); every compute call fails loudly with EllCudaError when no B200 is usable."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libellcuda.so")

# every symbol include/ellcuda.h declares
EXPORTS = [
    "ellcuda_batch_solve", "ellcuda_last_launch_count", "ellcuda_lowpass_table", "ellcuda_lowpass_band_edge",
    "ellcuda_create", "ellcuda_clone", "ellcuda_destroy", "ellcuda_xc", "ellcuda_set_xc", "ellcuda_tsq",
    "ellcuda_kappa", "ellcuda_get_mq", "ellcuda_dot_mv", "ellcuda_update", "ellcuda_update_begin", "ellcuda_update_finish",
    "ellcuda_grad_t_ptr", "ellcuda_exchange_export", "ellcuda_exchange_attach", "ellcuda_exchange_attach_ptrs",
    "ellcuda_exchange_block_ptr", "ellcuda_status_ptr", "ellcuda_last_status", "ellcuda_set_stream", "ellcuda_sync",
    "ellcuda_set_timing", "ellcuda_last_sweep_ms", "ellcuda_handle_launch_count",
    "ellcuda_math_probe", "ellcuda_last_error", "ellcuda_version", "ellcuda_device_count",
    "ellcuda_batch_solve_multi", "ellcuda_create_sharded", "ellcuda_group_size",
    "ellcuda_set_no_defer_trick", "ellcuda_set_use_parallel_cut", "ellcuda_seq_sum",
]


class EllCudaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libellcuda error {code}: {msg}")
        self.code = code


def build(force: bool = False) -> str:
    """compile csrc/*.cu for sm_100a (nvcc cross-compiles without a GPU)"""
    args = ["make", "-C", CSRC, "-j4"]
    if force:
        args.append("-B")
    p = subprocess.run(args, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("building libellcuda.so failed:\n" + p.stdout[-4000:] + p.stderr[-4000:])
    return LIB_PATH


_lib = None
_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:                      # a fresh checkout: compile once (nvcc is part of the image); never a CPU substitute
            build()
        except Exception as exc:
            raise EllCudaError(_abi.E_CUDA, f"{LIB_PATH} is missing and could not be built ({exc}); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.ellcuda_batch_solve.argtypes = [C.POINTER(_abi.BatchDesc), C.POINTER(_abi.BatchResult)]
    L.ellcuda_batch_solve.restype = C.c_int
    L.ellcuda_last_launch_count.restype = C.c_int64
    L.ellcuda_lowpass_table.argtypes = [C.c_int32, C.c_void_p]
    L.ellcuda_lowpass_table.restype = C.c_int
    L.ellcuda_lowpass_band_edge.argtypes = [C.c_int32, C.c_double]
    L.ellcuda_lowpass_band_edge.restype = C.c_int32
    L.ellcuda_last_error.restype = C.c_char_p
    L.ellcuda_math_probe.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32]
    L.ellcuda_math_probe.restype = C.c_int
    L.ellcuda_version.restype = C.c_int
    L.ellcuda_device_count.restype = C.c_int
    if hasattr(L, "ellcuda_create"):
        H = C.c_void_p
        L.ellcuda_create.argtypes = [C.POINTER(_abi.CreateDesc), C.POINTER(H)]
        L.ellcuda_create.restype = C.c_int
        L.ellcuda_create_sharded.argtypes = [C.POINTER(_abi.CreateDesc), C.c_int32, C.POINTER(C.c_int32), C.POINTER(H)]
        L.ellcuda_create_sharded.restype = C.c_int
        L.ellcuda_group_size.argtypes = [H]
        L.ellcuda_group_size.restype = C.c_int32
        for name in ("ellcuda_set_no_defer_trick", "ellcuda_set_use_parallel_cut"):
            getattr(L, name).argtypes = [H, C.c_int32]
            getattr(L, name).restype = C.c_int
        L.ellcuda_batch_solve_multi.argtypes = [C.POINTER(_abi.BatchDesc), C.POINTER(_abi.BatchResult), C.c_int32, C.POINTER(C.c_int32)]
        L.ellcuda_batch_solve_multi.restype = C.c_int
        L.ellcuda_clone.argtypes = [H, C.POINTER(H)]
        L.ellcuda_clone.restype = C.c_int
        L.ellcuda_destroy.argtypes = [H]
        L.ellcuda_destroy.restype = None
        for name in ("ellcuda_xc", "ellcuda_set_xc", "ellcuda_get_mq"):
            getattr(L, name).argtypes = [H, C.c_void_p]
            getattr(L, name).restype = C.c_int
        for name in ("ellcuda_tsq", "ellcuda_kappa"):
            getattr(L, name).argtypes = [H, _pd]
            getattr(L, name).restype = C.c_int
        L.ellcuda_dot_mv.argtypes = [H, C.c_void_p, C.c_void_p]
        L.ellcuda_dot_mv.restype = C.c_int
        L.ellcuda_seq_sum.argtypes = [H, C.c_void_p, C.c_int64, _pd]
        L.ellcuda_seq_sum.restype = C.c_int
        L.ellcuda_update.argtypes = [H, C.c_int32, C.c_void_p, C.c_double, C.c_int32, C.c_double, _pi]
        L.ellcuda_update.restype = C.c_int
        L.ellcuda_update_begin.argtypes = [H, C.c_int32, C.c_void_p, C.c_double, C.c_int32, C.c_double]
        L.ellcuda_update_begin.restype = C.c_int
        L.ellcuda_update_finish.argtypes = [H]
        L.ellcuda_update_finish.restype = C.c_int
        L.ellcuda_grad_t_ptr.argtypes = [H]
        L.ellcuda_grad_t_ptr.restype = C.c_void_p
        L.ellcuda_exchange_export.argtypes = [H, C.c_void_p]
        L.ellcuda_exchange_export.restype = C.c_int
        L.ellcuda_exchange_attach.argtypes = [H, C.c_int32, C.c_int32, C.c_void_p]
        L.ellcuda_exchange_attach.restype = C.c_int
        L.ellcuda_exchange_attach_ptrs.argtypes = [H, C.c_int32, C.c_int32, C.c_void_p]
        L.ellcuda_exchange_attach_ptrs.restype = C.c_int
        L.ellcuda_exchange_block_ptr.argtypes = [H]
        L.ellcuda_exchange_block_ptr.restype = C.c_void_p
        L.ellcuda_status_ptr.argtypes = [H]
        L.ellcuda_status_ptr.restype = C.c_void_p
        L.ellcuda_last_status.argtypes = [H, _pi]
        L.ellcuda_last_status.restype = C.c_int
        L.ellcuda_set_stream.argtypes = [H, C.c_void_p]
        L.ellcuda_set_stream.restype = C.c_int
        L.ellcuda_sync.argtypes = [H]
        L.ellcuda_sync.restype = C.c_int
        L.ellcuda_set_timing.argtypes = [H, C.c_int32]
        L.ellcuda_set_timing.restype = C.c_int
        L.ellcuda_last_sweep_ms.argtypes = [H]
        L.ellcuda_last_sweep_ms.restype = C.c_float
        L.ellcuda_handle_launch_count.argtypes = [H]
        L.ellcuda_handle_launch_count.restype = C.c_int64
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise EllCudaError(rc, lib().ellcuda_last_error().decode())
