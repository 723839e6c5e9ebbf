"""Batched mode: `count` independent cutting-plane problems solved entirely on the GPU (ellcuda_batch_solve)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from ._lib import check, lib
from .problem import BatchOutput, BatchProblem


def lowpass_table(n: int) -> np.ndarray:
    """LowpassOracle::new's spectrum (src/oracles/lowpass_oracle.rs:24-33), flattened (15n, n)."""
    t = np.zeros(15 * n * n)
    check(lib().ellcuda_lowpass_table(n, t.ctypes.data))
    return t


def solve_batch(prob: BatchProblem, out: BatchOutput | None = None, stream: int = 0) -> BatchOutput:
    """Host-buffer call: inputs are copied to the device, the persistent kernel runs the whole loop for every
    problem, results are copied back (all inside this call)."""
    if prob.on_device:
        raise ValueError("use solve_batch_device for device-resident problems")
    if out is None:
        out = BatchOutput.empty(prob.count, prob.n)
    d = prob.to_desc(stream)
    r = out.to_result()
    check(lib().ellcuda_batch_solve(C.byref(d), C.byref(r)))
    out.launches = lib().ellcuda_last_launch_count()
    return out


def solve_batch_multi(prob: BatchProblem, devices, out: BatchOutput | None = None) -> BatchOutput:
    """ellcuda_batch_solve_multi: the batch split into contiguous blocks over `devices` of this process (no launcher,
    no collective); host buffers."""
    if prob.on_device:
        raise ValueError("solve_batch_multi takes host arrays")
    if out is None:
        out = BatchOutput.empty(prob.count, prob.n)
    d = prob.to_desc(0)
    r = out.to_result()
    devs = (C.c_int32 * len(devices))(*devices)
    check(lib().ellcuda_batch_solve_multi(C.byref(d), C.byref(r), len(devices), devs))
    out.launches = lib().ellcuda_last_launch_count()
    return out


def solve_batch_device(prob: BatchProblem, res: dict, stream: int = 0) -> int:
    """Device-buffer call: every array of `prob` and every tensor in `res` (keys of ell_batch_result) is a CUDA
    tensor; the kernel is enqueued on `stream` and the call returns without synchronising.  Returns launches."""
    d = prob.to_desc(stream)
    d.on_device = 1
    r = _abi.BatchResult()
    for k, t in res.items():
        setattr(r, k, t.data_ptr())
    check(lib().ellcuda_batch_solve(C.byref(d), C.byref(r)))
    return lib().ellcuda_last_launch_count()
