"""Multi-GPU plumbing: one process per GPU, torch.distributed for the exchange.

  * batched mode shards PROBLEMS across ranks — independent units, no data-path collective; results are only
    gathered when the caller asks for them.
  * large-n mode shards ROWS of the one matrix: rank r owns rows [r·n/P, (r+1)·n/P) (held as the matching column
    block, M being symmetric).  Per update the only exchange is the allgather of the n entries of grad_t; omega,
    EllCalc, xc and kappa are then computed redundantly and identically on every rank (no second collective).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from . import _abi
from .problem import BatchOutput, BatchProblem


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous block [lo, hi) of `total` units for `rank`; sizes differ by at most one"""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def solve_batch_sharded(prob: BatchProblem, solver: Callable[[BatchProblem], BatchOutput], rank: int, world: int,
                        gather: bool = True, group=None) -> Optional[BatchOutput]:
    """Every rank holds the description of the whole batch, solves its block of problems with `solver`
    (batch.solve_batch on a GPU rank) and, if `gather`, rank 0 receives the concatenated result."""
    lo, hi = shard_range(prob.count, rank, world)
    local = solver(prob.slice(lo, hi)) if hi > lo else BatchOutput.empty(0, prob.n)
    if not gather or world == 1:
        return local
    import torch.distributed as dist
    parts = [None] * world if rank == 0 else None
    dist.gather_object(local, parts, dst=0, group=group)
    if rank != 0:
        return None
    out = parts[0]
    for p in parts[1:]:
        out = out.concat(p)
    return out


class ShardedEll:
    """Row-sharded `Ell` for large n: SearchSpace semantics of src/ell.rs with M split over the ranks of `group`.

    update(which, grad_dev, cut): sweep own slab (fused deferred rank-1 + GEMV) -> allgather(grad_t) -> finish.
    All ranks pass the same gradient (device tensor of n doubles) and end with identical xc / kappa / tsq.
    With exchange="p2p" (default) the allgather is fused into the kernel that completes the slab: it stores straight
    into every peer's exchange block over NVLink and `finish` waits on device-side flags."""

    def __init__(self, kappa, xc, *, mq=None, diag=None, no_defer_trick=False, use_parallel_cut=True, reduce_mode=0,
                 device=None, group=None, exchange="p2p", storage=0):
        import torch
        import torch.distributed as dist
        from .ell import Ell
        self._torch, self._dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        n = len(xc)
        self.n = n
        self.device = torch.cuda.current_device() if device is None else device
        self.lo, self.hi = shard_range(n, self.rank, self.world)
        if storage:   # symmetric storage: column-block pairs dealt round-robin to the ranks, the exchange is a reduction (p2p only)
            exchange = "p2p"
            self.ell = Ell(kappa, xc, mq=mq, diag=diag, device=self.device, no_defer_trick=no_defer_trick,
                           use_parallel_cut=use_parallel_cut, reduce_mode=reduce_mode, storage=storage,
                           shard=(self.rank, self.world) if self.world > 1 else None)
        else:
            self.ell = Ell(kappa, xc, mq=mq, diag=diag, device=self.device, no_defer_trick=no_defer_trick,
                           use_parallel_cut=use_parallel_cut, rows=(self.lo, self.hi), reduce_mode=reduce_mode)
        # run on torch's current stream so that the NCCL exchange is ordered with the sweeps; torch reports the legacy
        # default stream as handle 0, which the C ABI reads as "own stream" — cudaStreamLegacy (0x1) names it explicitly
        self.ell.set_stream(torch.cuda.current_stream(self.device).cuda_stream or 1)
        self.grad_t = torch.as_tensor(self.ell.grad_t_view(), device=f"cuda:{self.device}")
        self.even = n % self.world == 0
        # exchange of the grad_t slabs: "p2p" = stores into every peer's block from the producing kernel over NVLink
        # (CUDA IPC, no collective call per update); "nccl" = all_gather_into_tensor on torch's stream
        self.exchange_kind = exchange if self.world > 1 else "none"
        if self.exchange_kind == "p2p":
            handles = [None] * self.world
            dist.all_gather_object(handles, self.ell.exchange_export(), group=group)
            self.ell.exchange_attach(handles, self.rank)
            dist.barrier(group=group)

    def exchange(self):
        """allgather of grad_t slabs, in place (C1 of SURVEY.md §2)"""
        if self.exchange_kind != "nccl":
            return            # p2p: the begin kernels already pushed the slab to every peer
        dist = self._dist
        if self.even:
            dist.all_gather_into_tensor(self.grad_t, self.grad_t[self.lo:self.hi], group=self.group)
        else:
            for r in range(self.world):
                lo, hi = shard_range(self.n, r, self.world)
                dist.broadcast(self.grad_t[lo:hi], src=dist.get_global_rank(self.group, r) if self.group else r, group=self.group)

    def update_async(self, which: int, grad_dev, cut):
        self.ell.update_begin(which, grad_dev.data_ptr(), cut)
        self.exchange()
        self.ell.update_finish()

    def capture(self, which: int, grads, cut, stream):
        """CUDA graph of len(grads) consecutive updates (sweep -> exchange -> finish each), replayable any number of
        times: all per-update state, including the peer-exchange epoch, lives in device memory.  `stream` must be the
        torch stream this object was created under."""
        torch = self._torch
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for g in grads:
                self.update_async(which, g, cut)
        return graph

    def update(self, which: int, grad_dev, cut):
        self.update_async(which, grad_dev, cut)
        return self.ell.last_status()

    def xc(self) -> np.ndarray:
        return self.ell.xc()

    def tsq(self) -> float:
        return self.ell.tsq()

    @property
    def kappa(self) -> float:
        return self.ell.kappa
