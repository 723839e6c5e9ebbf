#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on B200: "ellipsoid iterations/sec (batched n=32) and Ell::update HBM GB/s at
n=16384".

  python bench.py [--gpus N] [--steps K] [--warmup W]            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                            the reference algorithm's CPU path (oracle port)

One JSON line on rank 0.  Headline workload (configs[2]): LowpassOracle n=32, 65536 independent problems per GPU,
Ell::new_with_scalar(40, 0), Options{50000, 1e-14}; a step = one ellcuda_batch_solve over the batch; the unit is one
ellipsoid iteration = one update_core call (oracle evaluation included in the time).  `large_n` carries the second
half of the metric (configs[4]): one n=16384 ellipsoid, ParallelCut(0.01, Some(0.10)) via update_bias_cut, 200
updates after 20 warm-up, row-sharded over the N ranks; the grad_t slabs are stored into every peer's memory over NVLink
by the sweep kernel itself (`--exchange nccl` selects an NCCL all_gather per update instead).  At N=1 the large-n leg runs
reduce_mode 0 (bit-identical, the headline) AND the tolerance modes the sharded runs use (2, 4, 8), so that scaling can be
read like for like; at every N it carries a `parity` record: the first two updates from M = I on this very configuration
against the CPU checker (tsq, kappa, all of xc, rank 0's slab of M).  `profit_1m` (configs[1]) and `lmi64` (configs[3]) are
timed the same way as the headline, each with its own CPU figure from the same box; `lowpass128` is LowpassOracle at the
dimension of the reference's stress test (128 taps) as a batch.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TAPS = 32
PER_GPU = 65536
LARGE_N = 16384
LARGE_UPDATES, LARGE_WARMUP = 200, 20


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def iterations_done(num_iters, exit_status):
    """update_core calls performed: the terminating iteration ran an update too (src/cutting_plane.rs:299-311)"""
    return int(num_iters.sum()) + int((exit_status != 5).sum())


def cpu_leg(make, sample, ni_gpu, es_gpu, stride):
    """the checker on a strided sample of the same batch: a CPU figure from this box and a parity check in one"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ellref_py as R
    T = R.hardware_threads()
    sp = make(sample)
    t0 = time.perf_counter()
    o = R.solve(sp, T)
    dt = time.perf_counter() - t0
    it = iterations_done(o.num_iters, o.exit_status)
    idx = np.arange(sample) * stride
    return {"value": it / dt, "unit": "iterations/s", "cores": T, "kind": "port",
            "sample": f"{sample} problems (every {stride}-th of the batch), {it} iterations, {dt:.1f} s, std::thread over problems",
            "sample_num_iters_and_status_identical": bool(np.array_equal(ni_gpu[idx], o.num_iters) and
                                                          np.array_equal(es_gpu[idx], o.exit_status))}

def strided(prob, stride, sample):
    """problems 0, stride, 2*stride, ... of a host batch"""
    sub = prob.slice(0, prob.count)
    for name, width in (("kappa0", 1), ("xc0", prob.n), ("diag0", prob.n), ("mq0", prob.n * prob.n), ("gamma0", 1)):
        a = getattr(prob, name)
        if a is not None and prob._stride(a, width):
            setattr(sub, name, np.ascontiguousarray(a.reshape(prob.count, -1)[::stride][:sample]))
    if prob.params is not None and prob.params.ndim == 2 and prob.params.shape[0] == prob.count:
        sub.params = np.ascontiguousarray(prob.params[::stride][:sample])
    sub.count = sample
    return sub


def load_counters():
    """per-kernel constants taken from the committed ncu captures (profiles/r02_counters.json, written by tools/ncu_counters.py)"""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_counters.json")) as f:
            return json.load(f)
    except Exception:
        return {}


COUNTERS = load_counters()


# ================================================================================================== reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ellref_py as R
    from ellalgo_rs_b200 import workloads as W
    T = R.hardware_threads()
    table = R.lowpass_table(N_TAPS)
    sample = max(16 * T, 32)
    prob = W.lowpass(sample, table, n=N_TAPS)
    for _ in range(args.warmup):
        R.solve(W.lowpass(max(T // 2, 1), table, n=N_TAPS), T)
    t_tot, it_tot = 0.0, 0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        o = R.solve(prob, T)
        t_tot += time.perf_counter() - t0
        it_tot += iterations_done(o.num_iters, o.exit_status)
    val = it_tot / t_tot
    desc = f"first {sample} instances of the lowpass n=32 batch per step, {T} std::threads over instances"
    line = {
        "impl": "reference", "metric": "ellipsoid iterations/sec (batched n=32)", "value": val, "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": "iterations/s", "cores": T, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the Rust reference cannot be built in this image (no cargo/rustc); this is oracle/ — a line-by-line C++ "
                "restatement pinned to the reference's own iteration counts — on the box's host cores",
    }
    print(json.dumps(line), flush=True)


def workload_config(gpus):
    return {"workload": "configs[2]: LowpassOracle n=32 (spectrum 480x32), 65536 independent problems per GPU, "
                        "Ell::new_with_scalar(40,0), Options{max_iters 50000, tol 1e-14}, cutting_plane_optim on-device",
            "problems_per_gpu": PER_GPU, "n": N_TAPS, "parallelism": f"problems sharded over {gpus} GPU(s), no collective",
            "l2": "L2 flushed (512 MiB memset) between timed steps; large_n streams a 2 GiB matrix (> L2)"}


# ================================================================================================== our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-large-n", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-1m", action="store_true")
    ap.add_argument("--skip-extras", action="store_true")
    ap.add_argument("--per-gpu", type=int, default=PER_GPU)
    ap.add_argument("--large-n", type=int, default=LARGE_N)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from ellalgo_rs_b200 import _abi as A
    from ellalgo_rs_b200 import batch
    from ellalgo_rs_b200 import workloads as W
    from ellalgo_rs_b200._lib import check, lib
    from ellalgo_rs_b200.cutting_plane import ParallelCut
    from ellalgo_rs_b200.parallel import ShardedEll
    from ellalgo_rs_b200.problem import BatchProblem

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)      # an explicit stream: handle 0 would mean "the library's own stream"
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    L = lib()
    peak, peak_src = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    # ---------------------------------------------------------------- batched lowpass n=32 (headline)
    cnt = args.per_gpu
    table = batch.lowpass_table(N_TAPS)
    host = W.lowpass(cnt, table, n=N_TAPS, first=rank * cnt)   # weak scaling: rank r owns global problems [r·cnt, (r+1)·cnt)
    host.device = local

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t

    pinned = {k: pin(getattr(host, k)) for k in ("kappa0", "xc0", "gamma0", "params", "table")}
    devin = {k: v.to(dev) for k, v in pinned.items()}
    dprob = BatchProblem(oracle_id=A.ORACLE_LOWPASS, n=N_TAPS, count=cnt, kappa0=devin["kappa0"], xc0=devin["xc0"],
                         gamma0=devin["gamma0"], params=devin["params"], table=devin["table"], aux=[15 * N_TAPS],
                         max_iters=host.max_iters, tolerance=host.tolerance, device=local)
    dres = {"xbest": torch.zeros(cnt, N_TAPS, dtype=torch.float64, device=dev),
            "has_xbest": torch.zeros(cnt, dtype=torch.uint8, device=dev),
            "num_iters": torch.zeros(cnt, dtype=torch.int64, device=dev),
            "exit_status": torch.zeros(cnt, dtype=torch.int32, device=dev),
            "gamma": torch.zeros(cnt, dtype=torch.float64, device=dev)}

    launches = 0
    for _ in range(args.warmup):
        batch.solve_batch_device(dprob, dres, sptr)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        flush_buf.zero_()
        ev[k][0].record(stream)
        launches += batch.solve_batch_device(dprob, dres, sptr)
        ev[k][1].record(stream)
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    t_dev = max_over_ranks(sum(step_ms) / 1e3)
    iters_step = iterations_done(dres["num_iters"].cpu().numpy(), dres["exit_status"].cpu().numpy())
    iters_all = sum_over_ranks(float(iters_step)) * args.steps
    value = iters_all / t_dev
    kern_ms = float(np.mean(step_ms))  # one launch per step: the step IS the kernel

    # e2e: host (pinned) buffers through the public C-ABI call, H2D + kernel + D2H inside the timed region
    hprob = BatchProblem(oracle_id=A.ORACLE_LOWPASS, n=N_TAPS, count=cnt, kappa0=pinned["kappa0"].numpy(),
                         xc0=pinned["xc0"].numpy(), gamma0=pinned["gamma0"].numpy(), params=pinned["params"].numpy(),
                         table=pinned["table"].numpy(), aux=[15 * N_TAPS], max_iters=host.max_iters,
                         tolerance=host.tolerance, device=local)
    hres_t = {"xbest": torch.zeros(cnt, N_TAPS, dtype=torch.float64).pin_memory(),
              "has_xbest": torch.zeros(cnt, dtype=torch.uint8).pin_memory(),
              "num_iters": torch.zeros(cnt, dtype=torch.int64).pin_memory(),
              "exit_status": torch.zeros(cnt, dtype=torch.int32).pin_memory(),
              "gamma": torch.zeros(cnt, dtype=torch.float64).pin_memory()}
    hd = hprob.to_desc(sptr)
    hr = A.BatchResult()
    for k, t in hres_t.items():
        setattr(hr, k, t.data_ptr())
    h2d = sum(int(pinned[k].numel()) * 8 for k in pinned)
    d2h = sum(int(t.numel()) * t.element_size() for t in hres_t.values())
    check(L.ellcuda_batch_solve(C.byref(hd), C.byref(hr)))  # warm
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_e2e = 0.0
    for k in range(args.steps):
        flush_buf.zero_()
        torch.cuda.synchronize()
        e0.record(stream)
        check(L.ellcuda_batch_solve(C.byref(hd), C.byref(hr)))
        e1.record(stream)
        e1.synchronize()
        t_e2e += e0.elapsed_time(e1) / 1e3
    barrier()
    t_e2e = max_over_ranks(t_e2e)
    e2e_iters = sum_over_ranks(float(iterations_done(hres_t["num_iters"].numpy(), hres_t["exit_status"].numpy()))) * args.steps
    e2e_value = e2e_iters / t_e2e
    clocks = sampler.stop() if sampler else None

    # ---------------------------------------------------------------- configs[0]: the reference's own CPU-runnable case (correctness)
    ex1_prob = W.example1()
    ex1_prob.device = local
    ex1 = batch.solve_batch(ex1_prob)
    example1 = {"workload": "configs[0]: src/example1.rs:39-50 through ellcuda_batch_solve", "num_iters": int(ex1.num_iters[0]),
                "reference_pin": 25, "exit_status": A.EXIT_NAMES.get(int(ex1.exit_status[0])), "xbest": [float(v) for v in ex1.xbest[0]],
                "equal_to_pin": bool(int(ex1.num_iters[0]) == 25)}

    # ---------------------------------------------------------------- north_star's own target shape: 2^20 n=32 problems / GPU
    extra_1m = None
    if not args.skip_1m:
        c1 = 1 << 20
        hb = W.bowl(c1, n=N_TAPS, first=rank * c1, max_iters=2000)     # SURVEY C3': cap 2000 -> every problem runs 2000 iterations
        db = BatchProblem(oracle_id=A.ORACLE_BOWL, n=N_TAPS, count=c1, kappa0=torch.from_numpy(hb.kappa0).to(dev),
                          xc0=torch.from_numpy(hb.xc0).to(dev), gamma0=torch.from_numpy(hb.gamma0).to(dev),
                          params=torch.from_numpy(hb.params).to(dev), max_iters=2000, tolerance=hb.tolerance, device=local)
        rb = {"num_iters": torch.zeros(c1, dtype=torch.int64, device=dev), "exit_status": torch.zeros(c1, dtype=torch.int32, device=dev),
              "xbest": torch.zeros(c1, N_TAPS, dtype=torch.float64, device=dev), "has_xbest": torch.zeros(c1, dtype=torch.uint8, device=dev),
              "gamma": torch.zeros(c1, dtype=torch.float64, device=dev)}
        batch.solve_batch_device(db, rb, sptr)
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(stream)
        batch.solve_batch_device(db, rb, sptr)
        b1.record(stream)
        barrier()
        tb = max_over_ranks(b0.elapsed_time(b1) / 1e3)
        itb = sum_over_ranks(float(iterations_done(rb["num_iters"].cpu().numpy(), rb["exit_status"].cpu().numpy())))
        extra_1m = {"workload": "SURVEY C3': 2^20 quadratic-bowl problems per GPU, n=32, Ell::new_with_scalar(100,0), max_iters capped at 2000",
                    "value": itb / tb, "unit": "iterations/s", "ms": 1e3 * tb,
                    "equivalent_streaming_GBs_per_gpu": itb / tb / world * 16.0 * N_TAPS * N_TAPS / 1e9,
                    "frac_of_nominal_8TBs": itb / tb / world * 16.0 * N_TAPS * N_TAPS / 8e12,
                    "target": "north_star: >= 50 % of the HBM roofline = 2.44e8 iterations/s per GPU"}
        del db, rb

    # ---------------------------------------------------------------- large n (second half of the metric)
    large = None
    if not args.skip_large_n:
        n = args.large_n
        grads = [torch.from_numpy(W.large_n_grad(n, k)).to(dev) for k in range(20)]
        cut = ParallelCut(*W.LARGE_N_CUT)
        bytes_update = 16.0 * n * n                       # read M once + write M once
        slab_bytes = bytes_update / world
        # kappa shrinks by ~1.5 % per update, so after ~290 of these cuts tsq falls under beta1^2 and the cut degrades to
        # a single/failed cut (cheaper sweeps).  Every phase therefore starts from a fresh M = I and stays below 260.

        def fresh(mode, warm=LARGE_WARMUP, storage=0):
            e = ShardedEll(1.0, np.zeros(n), reduce_mode=mode, device=local, exchange=args.exchange, storage=storage)
            for k in range(warm):
                e.update_async(A.UPDATE_BIAS, grads[k % 20], cut)
            return e

        def phases(mode, storage=0):
            """A: 200 updates as 10 replays of a 20-update CUDA graph; B: the sweep kernel alone (events around each launch)"""
            ell = fresh(mode, storage=storage)
            per_graph = 20
            graph = None if args.no_graph else ell.capture(A.UPDATE_BIAS, grads[:per_graph], cut, stream)
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = ell.ell.launches
            a0.record(stream)
            if graph is not None:
                for _ in range(LARGE_UPDATES // per_graph):
                    graph.replay()
            else:
                for k in range(LARGE_UPDATES):
                    ell.update_async(A.UPDATE_BIAS, grads[k % 20], cut)
            a1.record(stream)
            barrier()
            t_large = max_over_ranks(a0.elapsed_time(a1) / 1e3)
            per_update = 3 if storage else 2    # sym: sweep, combine (sharded: + push, reduce and broadcast of the slice), finish
            l_launch = (per_update * LARGE_UPDATES) if graph is not None else ell.ell.launches - l0
            status = int(ell.ell.last_status())
            kind = ell.exchange_kind
            del graph, ell
            ell = fresh(mode, storage=storage)
            ell.ell.set_timing(True)
            sweeps = []
            for k in range(60):
                ell.update_async(A.UPDATE_BIAS, grads[k % 20], cut)
                sweeps.append(ell.ell.last_sweep_ms())
            barrier()
            status_b = int(ell.ell.last_status())
            del ell
            return {"reduce_mode": mode, "updates_per_s": LARGE_UPDATES / t_large, "ms_per_update": 1e3 * t_large / LARGE_UPDATES,
                    "GBs": bytes_update * LARGE_UPDATES / t_large / 1e9, "sweep_ms": float(np.median(sweeps)),
                    "sweep_GBs_per_gpu": slab_bytes / (float(np.median(sweeps)) * 1e-3) / 1e9,
                    "last_status": status, "last_status_sweep_phase": status_b, "gpu_launches": int(l_launch), "exchange": kind}

        ref_cache = {}

        def parity(mode, storage=0):
            """the first two updates from M = I on THIS configuration (n, reduce_mode, exchange, world) against the CPU checker
            (oracle/, ~15 s per update on one host core, rank 0); every rank compares its own scalars and xc, rank 0 its slab"""
            ell = fresh(mode, warm=0, storage=storage)
            got = []
            for k in range(2):
                ell.update_async(A.UPDATE_BIAS, grads[k], cut)
                got.append((int(ell.ell.last_status()), ell.tsq(), ell.kappa, ell.xc()))
            whole = storage == 1
            slab = ell.ell.mq if (rank == 0 and not (whole and world > 1)) else None
            lo, hi = (0, n) if whole else (ell.lo, ell.hi)
            del ell
            ref, secs = [None], []
            if rank == 0:
                if "states" not in ref_cache:     # the checker's two updates are the same for every configuration: run them once
                    sys.path.insert(0, os.path.join(ROOT, "oracle"))
                    import ellref_py as R
                    mq, xc = np.eye(n), np.zeros(n)
                    kap, tsq = C.c_double(1.0), C.c_double(0.0)
                    states = []
                    for k in range(2):
                        g = W.large_n_grad(n, k)
                        t0 = time.perf_counter()
                        st = R.lib().ellref_ell_update_inplace(n, mq.ctypes.data, xc.ctypes.data, C.byref(kap), C.byref(tsq), 0, 1, 0,
                                                               g.ctypes.data, W.LARGE_N_CUT[0], 1, W.LARGE_N_CUT[1], None)
                        ref_cache.setdefault("secs", []).append(time.perf_counter() - t0)
                        states.append((int(st), tsq.value, kap.value, xc.copy()))
                    ref_cache["states"], ref_cache["mq"] = states, mq
                states, mq, secs = ref_cache["states"], ref_cache["mq"], ref_cache["secs"]
                ref = [states]
                if slab is not None:
                    den = float(np.abs(mq[lo:hi]).max())
                    m_err = float(np.abs(slab - mq[lo:hi]).max() / den)
                    m_same = bool(np.array_equal(slab, mq[lo:hi]))
                else:          # sharded symmetric storage: no rank holds rows; M is checked by tests/multi_gpu_check.py at small n
                    m_err, m_same = None, None
                del slab
            if world > 1:
                dist.broadcast_object_list(ref, src=0)
            worst, same, status_ok = 0.0, True, True
            for (st, tsq_g, kap_g, xc_g), (st_r, tsq_r, kap_r, xc_r) in zip(got, ref[0]):
                status_ok &= st == st_r
                same &= tsq_g == tsq_r and kap_g == kap_r and bool(np.array_equal(xc_g, xc_r))
                worst = max(worst, abs(tsq_g - tsq_r) / abs(tsq_r), abs(kap_g - kap_r) / abs(kap_r),
                            float(np.abs(xc_g - xc_r).max() / np.abs(xc_r).max()))
            worst = max_over_ranks(worst)
            same = max_over_ranks(0.0 if same else 1.0) == 0.0
            status_ok = max_over_ranks(0.0 if status_ok else 1.0) == 0.0
            rec = {"against": "oracle/ (CPU restatement of src/ell.rs:97-137), updates 1 and 2 from M = I, ParallelCut(0.01, 0.10)",
                   "n": n, "reduce_mode": mode, "storage": "symmetric" if storage else "full", "world": world, "status_equal": status_ok,
                   "max_rel_err_tsq_kappa_xc_over_ranks": worst, "bit_identical_tsq_kappa_xc": same, "tolerance": 1e-9}
            if rank == 0:
                rec.update({"rank0_slab_of_M_max_rel_err": m_err, "rank0_slab_of_M_bit_identical": m_same,
                            "checker_s_per_update": secs})
                exact = mode == 0 and storage == 0      # full storage in reduce_mode 0 must be bit-identical
                ok = status_ok and worst <= 1e-9 and (m_err is None or m_err <= 1e-9) and (not exact or (same and m_same))
                rec["pass"] = bool(ok)
            return rec, secs

        mode = 0 if world == 1 else world    # k row segments keep 256 sweep warps per GPU when the slab has n/P columns
        main_phase = phases(mode)
        by_mode = {str(mode): main_phase}
        if world == 1:
            for m in (2, 4, 8):              # the modes the sharded runs use, on one GPU: the like-for-like baselines
                by_mode[str(m)] = phases(m)
        par, ref_secs = parity(mode)
        parities = {str(mode): par}
        if world == 1 and not args.skip_cpu:
            parities["8"], _ = parity(8)
        symmetric = None
        if True:
            # tolerance-mode option: only the lower block-triangle is stored and streamed (8*n^2 bytes per update); sharded, the
            # column-block pairs are dealt round-robin to the ranks and the exchange is a reduction of flagged lines over NVLink
            sp = phases(0, storage=1)
            real = 8.0 * n * n + 8.0 * n * 64            # lower block-triangle incl. full diagonal tiles, read + write
            symmetric = {"what": "ELL_STORAGE_SYMMETRIC: lower block-triangle of 64x64 tiles only; sums regrouped, results to 1e-9",
                         "updates_per_s": sp["updates_per_s"], "ms_per_update": sp["ms_per_update"], "sweep_ms": sp["sweep_ms"],
                         "bytes_per_update": real, "GBs": real * sp["updates_per_s"] / 1e9,
                         "sweep_GBs_per_gpu": real / world / (sp["sweep_ms"] * 1e-3) / 1e9,
                         "sweep_frac_of_measured_peak": real / world / (sp["sweep_ms"] * 1e-3) / 1e9 / peak,
                         "last_status": sp["last_status"], "gpu_launches": sp["gpu_launches"]}
            if not args.skip_cpu:
                symmetric["parity"], _ = parity(0, storage=1)
        sweep_ms = main_phase["sweep_ms"]
        large = {
            "metric": "Ell::update HBM GB/s at n=16384 (whole job, 16*n^2 algorithmic bytes per update)",
            "value": main_phase["GBs"], "unit": "GB/s", "updates_per_s": main_phase["updates_per_s"],
            "ms_per_update": main_phase["ms_per_update"],
            "n": n, "updates": LARGE_UPDATES, "warmup": LARGE_WARMUP, "scaling": "strong", "reduce_mode": mode,
            "last_status": main_phase["last_status"], "last_status_sweep_phase": main_phase["last_status_sweep_phase"],
            "gpu_launches": main_phase["gpu_launches"], "cuda_graph": not args.no_graph,
            "by_reduce_mode": by_mode,
            "symmetric_storage": symmetric,
            "like_for_like": "N>1 runs reduce_mode=N (N ordered row segments per column block, 1e-9 gate); compare with "
                             "by_reduce_mode[str(N)] of the N=1 line, not with its bit-identical reduce_mode 0 headline",
            "parity": parities,
            "collective": None if world == 1 else f"{main_phase['exchange']}: grad_t slabs (n*8 bytes in total) " + (
                "stored into every peer's block over NVLink by the kernel that completes them; finish waits on device flags"
                if main_phase["exchange"] == "p2p" else "NCCL all_gather_into_tensor per update"),
            "roofline": {"bound": "hbm", "kernel": "sweep_kernel (fused deferred rank-1 + GEMV, TMA-staged)",
                         "achieved": slab_bytes / (sweep_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": slab_bytes / (sweep_ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                         "frac_of_nominal_8TBs": slab_bytes / (sweep_ms * 1e-3) / 1e9 / 8000.0,
                         "ms_per_launch": sweep_ms, "algorithmic_bytes_per_launch": slab_bytes,
                         "traffic": COUNTERS.get("sweep_n16384", {}).get("dram_bytes_per_launch") if (world == 1 and n == 16384) else None,
                         "traffic_source": COUNTERS.get("sweep_n16384", {}).get("source") if (world == 1 and n == 16384) else None,
                         "timing": "median of 60 CUDA-event pairs around the sweep launches of an eager run (phase B)"},
        }
        ref_cache.pop("mq", None)   # 2 GiB of host memory
        if rank == 0 and world == 1 and ref_secs:
            dt = ref_secs[0]
            large["cpu_baseline"] = {"value": 16.0 * n * n / dt / 1e9, "unit": "GB/s", "s_per_update": dt, "cores": 1, "kind": "port",
                                     "sample": "1 update of the reference algorithm (single-threaded, as the Rust code is) — the first "
                                               "update of the parity record"}
        # phase C — e2e: SearchSpace::update_bias_cut with a HOST gradient per call (H2D 8n bytes, D2H status), whole matrix
        if world == 1:
            ell = fresh(0)
            hg = [torch.from_numpy(W.large_n_grad(n, k)).pin_memory() for k in range(20)]
            st = C.c_int32()
            K2 = 100
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(K2):
                check(L.ellcuda_update(ell.ell._h, A.UPDATE_BIAS, hg[k % 20].data_ptr(), cut.beta0, 1, cut.beta1, C.byref(st)))
            t1 = time.perf_counter() - t0
            large["e2e"] = {"value": bytes_update * K2 / t1 / 1e9, "unit": "GB/s", "updates_per_s": K2 / t1,
                            "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 4, "updates": K2, "last_status": int(st.value),
                            "timing": "host wall clock around synchronous ellcuda_update calls"}
            del ell

    # ---------------------------------------------------------------- configs[1] and configs[3], timed like the headline
    extras = {}
    if not args.skip_extras:
        def timed_batch(host, keys, reps=2):
            dv = {k: torch.from_numpy(np.ascontiguousarray(getattr(host, k))).to(dev) for k in keys if getattr(host, k) is not None}
            dp = BatchProblem(oracle_id=host.oracle_id, n=host.n, count=host.count, mode=host.mode, kind=host.kind,
                              max_iters=host.max_iters, tolerance=host.tolerance, aux=list(host.aux), device=local, **dv)
            res = {"xbest": torch.zeros(host.count, host.n, dtype=torch.float64, device=dev),
                   "has_xbest": torch.zeros(host.count, dtype=torch.uint8, device=dev),
                   "num_iters": torch.zeros(host.count, dtype=torch.int64, device=dev),
                   "exit_status": torch.zeros(host.count, dtype=torch.int32, device=dev),
                   "gamma": torch.zeros(host.count, dtype=torch.float64, device=dev)}
            batch.solve_batch_device(dp, res, sptr)     # warm
            barrier()
            best = None
            for _ in range(reps):
                flush_buf.zero_()
                b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                b0.record(stream)
                batch.solve_batch_device(dp, res, sptr)
                b1.record(stream)
                barrier()
                t = max_over_ranks(b0.elapsed_time(b1) / 1e3)
                best = t if best is None else min(best, t)
            ni, es = res["num_iters"].cpu().numpy(), res["exit_status"].cpu().numpy()
            its = sum_over_ranks(float(iterations_done(ni, es)))
            # e2e: host buffers through ellcuda_batch_solve (H2D + kernel + D2H inside the call)
            host.device = local
            t_e2e = None
            for _ in range(2):          # the first call also pays the allocations of its staging buffers
                t0 = time.perf_counter()
                ho = batch.solve_batch(host, stream=sptr)
                dt = time.perf_counter() - t0
                t_e2e = dt if t_e2e is None else min(t_e2e, dt)
            t_e2e = max_over_ranks(t_e2e)
            its_e2e = sum_over_ranks(float(iterations_done(ho.num_iters, ho.exit_status)))
            return {"value": its / best, "unit": "iterations/s", "ms": 1e3 * best, "problems_per_gpu": host.count,
                    "problems_per_s": host.count * world / best, "iterations_per_gpu": iterations_done(ni, es),
                    "e2e": {"value": its_e2e / t_e2e, "unit": "iterations/s", "timing": "host wall clock around ellcuda_batch_solve with host buffers (pageable), best of 2"}}, ni, es

        cpu_ok = rank == 0 and world == 1 and not args.skip_cpu
        c1 = 1 << 20
        hp = W.profit(c1, first=rank * c1)
        rec, ni, es = timed_batch(hp, ("kappa0", "xc0", "diag0", "gamma0", "params"))
        rec["workload"] = "configs[1]: ProfitOracle (Cobb-Douglas, n=2), 2^20 independent instances per GPU, Ell::new([100,100],[0,0]), default Options"
        if cpu_ok:
            rec["cpu_baseline"] = cpu_leg(lambda k: strided(hp, 1, k), c1, ni, es, 1)     # the whole batch: < 1 s on 16 threads
        extras["profit_1m"] = rec
        for name, general, cnt_l in (("lmi64_diag", False, 1 << 15), ("lmi64_general", True, 1 << 15)):
            hl = W.lmi(cnt_l, n=64, m=16, first=rank * cnt_l, general=general)
            keys = ("kappa0", "xc0", "mq0", "gamma0", "table")
            rec, ni, es = timed_batch(hl, keys)
            rec["workload"] = ("configs[3]: LMIOracle (ldlt_mgr) feasibility, n=64, m=16, EllStable, 2^15 instances per GPU; " +
                               ("C4b EllStable::new_with_matrix (general LDL^T recurrence)" if general else
                                "C4a EllStable::new_with_scalar (the factor provably stays diagonal)"))
            if cpu_ok:
                st_, sm_ = (1, cnt_l) if general else (4, cnt_l // 4)     # C4b runs end after a few iterations: take them all
                rec["cpu_baseline"] = cpu_leg(lambda k, hl=hl, st_=st_: strided(hl, st_, k), sm_, ni, es, st_)
            extras[name] = rec
        # the reference's own stress dimension (tests/stress_tests.rs:6-14: LowpassOracle with 128 taps) as a batch: one CTA per
        # problem, the packed lower triangle of M in shared memory (batch_ellcta.cu)
        n_l, cnt_l = 128, 1776
        hs = W.lowpass(cnt_l, batch.lowpass_table(n_l), n=n_l, first=rank * cnt_l, max_iters=3000)
        rec, ni, es = timed_batch(hs, ("kappa0", "xc0", "gamma0", "params", "table"))
        rec["workload"] = ("LowpassOracle with 128 taps (the dimension of tests/stress_tests.rs:6-14), Ell::new_with_scalar(40, 0), sane bounds, "
                           "1776 instances per GPU, Options{max_iters 3000, tol 1e-14}")
        if cpu_ok:
            rec["cpu_baseline"] = cpu_leg(lambda k: strided(hs, 16, k), cnt_l // 16, ni, es, 16)
        extras["lowpass128"] = rec

    # ---------------------------------------------------------------- CPU baseline of the headline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ellref_py as R
        T = R.hardware_threads()
        sample = min(max(256 * T, 512), cnt)    # ≈10 s of CPU work on 16 threads
        stride = max(cnt // sample, 1)          # spread over the whole batch, so the parity check is not just its head
        cpu = cpu_leg(lambda k: strided(host, stride, k), sample, dres["num_iters"].cpu().numpy(),
                      dres["exit_status"].cpu().numpy(), stride)
        cpu["per_core"] = cpu["value"] / cpu["cores"]

    if rank == 0:
        alg_bytes = iters_step * 16.0 * N_TAPS * N_TAPS
        ctr = COUNTERS.get("lowpass32", {})
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        clk_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        # FP64 pipe: 16 lanes per SM sub-partition per clock = one warp instruction every 2 clocks; the reference's arithmetic
        # is unfused (Rust never contracts a*b+c, the library is built -fmad=false), so an instruction is 32 flops, not 64
        fp64_peak = sm_count * 4 * 16 * clk_mhz * 1e6 / 1e12
        per_iter = ctr.get("fp64_warp_instr_per_iteration")
        fp64_ach = None if per_iter is None else per_iter * 32.0 * iters_step / (kern_ms * 1e-3) / 1e12
        # shared-memory data pipe: one 128-byte wavefront per SM per clock; the kernel's wavefronts per iteration are a property
        # of the code (ncu, committed), the iterations and the time are measured in this run
        wf_iter = ctr.get("shared_wavefronts_per_iteration")
        smem_peak = sm_count * 128.0 * clk_mhz * 1e6 / 1e9
        smem_ach = None if wf_iter is None else wf_iter * 128.0 * iters_step / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": "ellipsoid iterations/sec (batched n=32)", "value": value, "unit": "iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world),
            "iterations_per_step_per_gpu": iters_step,
            "roofline": {"bound": "shared_memory_pipe",
                         "kernel": "lowpass32_tmem_kernel (persistent, one warp per problem, M in tensor memory)",
                         "achieved": smem_ach, "peak": smem_peak, "unit": "GB/s",
                         "frac": None if smem_ach is None else smem_ach / smem_peak,
                         "peak_source": f"{sm_count} SMs x 128 B per clock (one shared-memory wavefront per SM per clock) x {clk_mhz:.0f} MHz "
                                        "sampled under load; tools/ubench_tmem.cu measures the same 128 B/clk/SM for LDS.64",
                         "shared_wavefronts_per_iteration": wf_iter, "bytes_per_launch": None if wf_iter is None else wf_iter * 128.0 * iters_step,
                         "why_this_bound": "highest-utilised unit of the kernel (ncu: shared-memory data pipe 83 %, FP64 pipe 52 %, issue "
                                           "slots 59 %); M lives in tensor memory and x in registers, so what is left in shared memory is the "
                                           "spectrum (290 of the 468 wavefronts per iteration: 145 rows x 256 B) and the broadcast vectors",
                         "ms_per_launch": kern_ms, "counters_source": ctr.get("source"), "ncu_pipe_utilisation": ctr.get("pipes"),
                         "fp64_pipe": {"achieved": fp64_ach, "peak": fp64_peak, "unit": "TFLOP/s",
                                       "frac": None if fp64_ach is None else fp64_ach / fp64_peak,
                                       "fp64_warp_instr_per_iteration": per_iter,
                                       "peak_source": f"{sm_count} SMs x 4 sub-partitions x 16 FP64 lanes/clk x {clk_mhz:.0f} MHz, unfused DMUL/DADD "
                                                      "= 1 flop per lane (the path may not contract to FMA); tools/ubench_fp64_issue.cu: 2 clk per "
                                                      "warp instruction"},
                         "traffic": ctr.get("dram_bytes_per_launch"), "traffic_source": ctr.get("source"),
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu capture at this launch shape (65536 problems): "
                                         "inputs and results only, M never leaves the SM (tensor memory + registers)",
                         "equivalent_streaming": {"achieved": alg_bytes / (kern_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                                  "frac": alg_bytes / (kern_ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                                                  "algorithmic_bytes_per_launch": alg_bytes,
                                                  "note": "SURVEY.md 8d / north_star's definition: 16*n^2 bytes per iteration as if M "
                                                          "were streamed from HBM once per update; not a physical bandwidth"}},
            "e2e": {"value": e2e_value, "unit": "iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * t_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "cpu_baseline": cpu,
            "example1": example1,
            "bowl_1m": extra_1m,
            "profit_1m": extras.get("profit_1m"), "lmi64_diag": extras.get("lmi64_diag"), "lmi64_general": extras.get("lmi64_general"),
            "lowpass128": extras.get("lowpass128"),
            "large_n": large,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
