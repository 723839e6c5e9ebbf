// stream_update.cu — large-n mode: ONE ellipsoid whose n×n matrix lives in HBM (src/ell.rs:9-16, 97-137).
//
// The reference streams M three times per update (GEMV read, rank-1 read + write).  Here one update costs ONE read
// and ONE write of M (16·n² bytes, the algorithmic minimum), by deferring the rank-1 of update k into the sweep
// that forms grad_t of update k+1:
//
//     sweep(k+1):   M <- (M - (ratio_k·g̃_k[hi])·g̃_k[lo]) [·kappa if no_defer_trick]     src/ell.rs:117-128, 132-135
//                   g̃_{k+1}[c] = sum_i M[i][c]·g_{k+1}[i]   (i ascending)                src/arr.rs:426-442
//     finish(k+1):  omega, tsq, EllCalc, xc, kappa, and {g̃, ratio·g̃} for the next sweep    src/ell.rs:103-115, 119, 130
//
// Two kernels per update.  The sweep's tail also does the ordered combine of row-segment partials and, for
// row-sharded handles, pushes the finished block of g̃ into every peer GPU's memory over NVLink.
//
// M is symmetric bit for bit (the reference mirrors every write), so the column sum over i is the reference's row
// sum over j, term for term and in the same order; with reduce_mode 0 grad_t and omega are bit-identical to Rust.
//
// Kernel shape: a CTA is ONE WARP that owns 64 adjacent columns (2 per lane) and walks down all rows; tiles of
// 64 rows × 64 columns (32 KB) are staged in shared memory by TMA (cp.async.bulk.tensor.2d, 3-stage mbarrier ring),
// updated in place, and written back by TMA.  Global traffic is therefore fully coalesced 512-byte row segments
// issued by the copy engine; the lanes only touch shared memory (conflict-free 16-byte accesses).  n=16384 gives
// 256 such warps (all resident, 2–3 per SM); the sequential accumulation a lane carries costs ~20 cycles per row
// and hides under the HBM time of the tile.  With row sharding a rank holds an n × w column block of M (= its w
// rows, by symmetry) and reduce_mode k>0 splits the rows into k ordered segments so that enough warps exist.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"
#include "stream_common.cuh"

namespace ellcuda {
namespace {

struct SweepArgs {
    const double* g;           // gradient of this update, readable for npad rows (user buffer if n % 64 == 0, else a padded copy)
    const double2* pend;       // {grad_t[i], ratio·grad_t[i]} of the deferred rank-1 (written by finish), padded with zeros
    const Scal* sc;
    double* partial;           // nseg × ld_part column sums per row segment
    unsigned* cnt_block;       // arrivals per column block (self-resetting)
    double* own_blk;           // this rank's exchange block
    double* out;               // non-null: plain GEMV — column sums go here (local column index), nothing else is published
    long long ld_part, n, w, c_glob0;
    int nseg;
    int exchange;              // 0: flush only (no grad_t is produced); 1: produce grad_t (+ push to peers if attached)
    PeerTable pt;
};

// a padded copy of the gradient, only needed when n is not a multiple of the tile height (bulk loads must stay in bounds)
__global__ void copy_pad_kernel(const double* __restrict__ g_in, double* __restrict__ g_out, long long n, long long npad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npad) g_out[i] = i < n ? g_in[i] : 0.0;
}

// shared-memory ring of one sweep warp: [kStages tiles][kStages pend rows][kStages g rows][kStages mbarriers]
template <int kCols, int kRows, int kStages>
struct Ring {
    static constexpr int kTileBytes = kCols * kRows * 8;
    static constexpr int kPendBytes = kRows * 16;
    static constexpr int kGBytes = kRows * 8;
    unsigned char *tiles, *pends, *gss;
    unsigned long long* bars;
    __device__ explicit Ring(unsigned char* smem_raw) {
        tiles = smem_raw;
        pends = smem_raw + kStages * kTileBytes;
        gss = pends + kStages * kPendBytes;
        bars = reinterpret_cast<unsigned long long*>(gss + kStages * kGBytes);
    }
    // a stage's barrier takes TWO arrivals per phase: the matrix tile and the vectors.  The tile of the first stages is
    // requested before pdl_wait() (M is not written by the preceding finish kernel), their vectors after it.
    __device__ __forceinline__ void issue_m(const CUtensorMap* map, int t, long long r0, long long c_loc0) const {
        const int s = t % kStages;
        mbar_expect_tx(&bars[s], kTileBytes);
        tma_load_2d(tiles + s * kTileBytes, map, (int)c_loc0, (int)r0, &bars[s]);
    }
    template <bool PENDING>
    __device__ __forceinline__ void issue_v(const SweepArgs& a, int t, long long r0) const {
        const int s = t % kStages;
        mbar_expect_tx(&bars[s], kGBytes + (PENDING ? kPendBytes : 0));
        bulk_load_1d(gss + s * kGBytes, a.g + r0, kGBytes, &bars[s]);
        if (PENDING) bulk_load_1d(pends + s * kPendBytes, a.pend + r0, kPendBytes, &bars[s]);
    }
};

// kW = warps per CTA.  One warp (two columns per lane) is the default.  ncu shows that warp issuing 18 % of the time with the
// rest spent on its own dependencies (2.2 us of instruction latency per 32 KB tile against 2.56 us of HBM time), so a
// two-warp form was built (warp w: columns 32w..32w+31, one per lane; still per column, rows ascending: bit-identical) —
// it measured no faster (see launch_sweep), the per-tile CTA barrier costs what the shorter instruction stream saves.
template <int kW>
__device__ __forceinline__ void cta_sync() {
    if (kW == 1) __syncwarp();
    else __syncthreads();
}

template <int kCols, int kRows, int kStages, int kW, bool PENDING, bool SCALE>
__device__ __forceinline__ void sweep_body(const CUtensorMap* map, const SweepArgs& a, long long row_lo, long long row_hi, double scale,
                                           const Ring<kCols, kRows, kStages>& ring, int ntiles, int pre, double& acc0, double& acc1) {
    constexpr int kTileBytes = kCols * kRows * 8;
    constexpr int kPendBytes = kRows * 16;
    constexpr int kGBytes = kRows * 8;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long c_loc0 = (long long)blockIdx.x * kCols;  // first local column of this CTA's block
    if (threadIdx.x == 0)
        for (int t = 0; t < pre; ++t) ring.template issue_v<PENDING>(a, t, row_lo + (long long)t * kRows);   // their tiles are on the way

    // this lane's column(s) (global indices) and their entries of the deferred rank-1 vectors
    constexpr int kPerLane = (kCols == 64 && kW == 1) ? 2 : 1;
    const int lcol = kW == 2 ? 32 * wid + lane : kPerLane * (kCols < 32 ? (lane % kCols) : lane);   // column inside the block
    const long long cg0 = a.c_glob0 + c_loc0 + lcol;
    double gt_c0 = 0.0, gt_c1 = 0.0, rg_c0 = 0.0, rg_c1 = 0.0;
    if (PENDING) {
        const double2 p0 = a.pend[cg0], p1 = a.pend[cg0 + 1];  // pend is padded, always readable
        gt_c0 = p0.x; rg_c0 = p0.y;
        gt_c1 = p1.x; rg_c1 = p1.y;
    }
    const long long blk_lo = a.c_glob0 + c_loc0, blk_hi = blk_lo + kCols;  // global column range of the block

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % kStages;
        mbar_wait(&ring.bars[s], (unsigned)((t / kStages) & 1));
        double* tile = reinterpret_cast<double*>(ring.tiles + s * kTileBytes);
        const double2* pd = reinterpret_cast<const double2*>(ring.pends + s * kPendBytes);
        const double* gs = reinterpret_cast<const double*>(ring.gss + s * kGBytes);
        const long long r0 = row_lo + (long long)t * kRows;
        const int nrows = (int)((row_hi - r0) < kRows ? (row_hi - r0) : kRows);
        double* base = kW == 2 ? tile + 32 * wid : tile;
        if (r0 + kRows <= blk_lo)
            tile_rows<kCols, 0, PENDING, SCALE, kCols, kPerLane>(base, gs, pd, lane, nrows, 0, 0, gt_c0, gt_c1, rg_c0, rg_c1, scale, acc0, acc1);
        else if (r0 >= blk_hi - 1)
            tile_rows<kCols, 1, PENDING, SCALE, kCols, kPerLane>(base, gs, pd, lane, nrows, 0, 0, gt_c0, gt_c1, rg_c0, rg_c1, scale, acc0, acc1);
        else
            tile_rows<kCols, 2, PENDING, SCALE, kCols, kPerLane>(base, gs, pd, lane, nrows, (int)(r0 - blk_lo), kW == 2 ? lcol : kPerLane * lane,
                                                                gt_c0, gt_c1, rg_c0, rg_c1, scale, acc0, acc1);
        if (PENDING) fence_async_smem();  // generic-proxy writes -> visible to the TMA store
        cta_sync<kW>();
        if (threadIdx.x == 0) {
            if (PENDING) {
                tma_store_2d(map, (int)c_loc0, (int)r0, tile);
                bulk_commit();
            }
            const int nt = t + kStages - 1;  // refill the stage whose store was issued one iteration ago
            if (nt < ntiles) {
                if (PENDING) bulk_wait_read<1>();
                const long long rn = row_lo + (long long)nt * kRows;
                ring.issue_m(map, nt, rn, c_loc0);
                ring.template issue_v<PENDING>(a, nt, rn);
            }
        }
        if (kW == 1) __syncwarp();
    }
    if (PENDING && threadIdx.x == 0) bulk_wait_all();
}

// grid = (column blocks, row segments); block = 32·kW threads.
template <int kCols, int kRows, int kStages, int kW>
__global__ void __launch_bounds__(32 * kW) sweep_kernel(const __grid_constant__ CUtensorMap map, const SweepArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ELL_STATIC_SHARED unsigned s_prev;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int cb = blockIdx.x, seg = blockIdx.y;
    const long long rows_per = ((a.n + a.nseg - 1) / a.nseg + kRows - 1) / kRows * kRows;
    const long long row_lo = (long long)seg * rows_per;
    long long row_hi = row_lo + rows_per;
    if (row_hi > a.n) row_hi = a.n;
    const Ring<kCols, kRows, kStages> ring(smem_raw);
    const int ntiles = row_lo < row_hi ? (int)((row_hi - row_lo + kRows - 1) / kRows) : 0;
    const int pre = ntiles < (kStages - 1) ? ntiles : (kStages - 1);
    // ---- prologue, may overlap the preceding finish kernel (PDL): barriers, and the first tiles of M are requested.
    // M is written only by sweeps of this handle, and the kernel before this one is never such a sweep still in flight
    // (a finish kernel triggers its successor only after its own pdl_wait, i.e. after the previous sweep completed).
    if (ntiles > 0) {
        if (threadIdx.x == 0) {
            for (int s = 0; s < kStages; ++s) mbar_init(&ring.bars[s], 2);
            mbar_fence_init();
            for (int t = 0; t < pre; ++t) ring.issue_m(&map, t, row_lo + (long long)t * kRows, (long long)cb * kCols);
        }
        cta_sync<kW>();
    }
    if (a.exchange) pdl_trigger();   // the finish kernel's CTAs may become resident now (a flush sweep keeps its successor out)
    pdl_wait();                      // the preceding kernel (finish of the previous update, or whatever produced g) is complete
    double acc0 = 0.0, acc1 = 0.0;   // an empty segment contributes exact zeros
    if (ntiles > 0) {
        const int pending = a.sc->pending;
        const double scale = a.sc->scale;
        if (!pending) sweep_body<kCols, kRows, kStages, kW, false, false>(&map, a, row_lo, row_hi, 1.0, ring, ntiles, pre, acc0, acc1);
        else if (scale == 1.0) sweep_body<kCols, kRows, kStages, kW, true, false>(&map, a, row_lo, row_hi, 1.0, ring, ntiles, pre, acc0, acc1);
        else sweep_body<kCols, kRows, kStages, kW, true, true>(&map, a, row_lo, row_hi, scale, ring, ntiles, pre, acc0, acc1);
    }
    if (!a.exchange) return;   // flush: only the deferred rank-1 was wanted

    // ---- tail: the last row segment of a column block sums the segments in order and publishes the block
    constexpr int kPerLane = (kCols == 64 && kW == 1) ? 2 : 1;
    const bool lane_on = kCols >= 32 || lane < kCols;
    const long long cl = (long long)cb * kCols + (kW == 2 ? 32 * wid + lane : kPerLane * (lane_on ? lane : 0));   // local column of this lane
    if (a.nseg > 1) {
        double* out = a.partial + (long long)seg * a.ld_part;
        if (lane_on) {
            out[cl] = acc0;          // ld_part is a multiple of the block width: always in range
            if (kPerLane == 2) out[cl + 1] = acc1;
        }
        __threadfence();
        cta_sync<kW>();
        if (threadIdx.x == 0) s_prev = atomicAdd(&a.cnt_block[cb], 1u);
        cta_sync<kW>();
        const unsigned prev = s_prev;
        if (prev != (unsigned)(a.nseg - 1)) return;
        if (threadIdx.x == 0) a.cnt_block[cb] = 0;
        __threadfence();
        acc0 = __ldcg(a.partial + cl);
        acc1 = kPerLane == 2 ? __ldcg(a.partial + cl + 1) : 0.0;
        for (int s = 1; s < a.nseg; ++s) {   // segment order: deterministic
            acc0 += __ldcg(a.partial + (long long)s * a.ld_part + cl);
            if (kPerLane == 2) acc1 += __ldcg(a.partial + (long long)s * a.ld_part + cl + 1);
        }
    }
    const bool v0 = lane_on && cl < a.w, v1 = kPerLane == 2 && cl + 1 < a.w;
    if (a.out) {   // Arr::dot_mv for its own sake (power iteration, conjugate gradient): no exchange, no state
        if (v0) a.out[cl] = acc0;
        if (v1) a.out[cl + 1] = acc1;
        return;
    }
    if (a.pt.nranks == 0) {   // no peer exchange: plain doubles (read by finish after this kernel, or gathered by the caller)
        if (v0) a.own_blk[a.c_glob0 + cl] = acc0;
        if (v1) a.own_blk[a.c_glob0 + cl + 1] = acc1;
        return;
    }
    if (a.sc->failed) return;   // sticky peer failure: publish nothing, the peers' waits expire too
    const unsigned long long epoch = a.sc->epoch;
    const unsigned flag = (unsigned)epoch;
    const long long c = a.c_glob0 + cl;
    for (int r = 0; r < a.pt.nranks; ++r) {   // this block of grad_t as flagged lines, into every rank's buffer (NVLink stores)
        uint4* lines = xblk_lines(a.pt.blk[r], a.n, (int)(epoch & 1ULL));
        if (v0) ll_store(lines + c, acc0, flag);
        if (v1) ll_store(lines + c + 1, acc1, flag);
    }
}

// omega, tsq, EllCalc, xc, kappa, and the vectors of the next deferred rank-1 (src/ell.rs:103-115, 119, 130-135).
// Identical on every rank.  A few CTAs: each one derives the scalars redundantly (same inputs, same order, same bits)
// and updates its slice of xc / pend; the CTA that arrives last — when every other CTA has read `sc` — writes the new
// scalars back, so no CTA can observe a half-updated state.  With peer exchange grad_t is read from this rank's flagged
// lines (polling until every slab has arrived), otherwise from the plain buffer the sweep (or the caller's allgather) filled.
#ifndef ELLCUDA_FIN_THREADS
#define ELLCUDA_FIN_THREADS 1024   // the host emulation under ThreadSanitizer builds with 128 (it tracks <= 256 running threads)
#endif
constexpr int kFinThreads = ELLCUDA_FIN_THREADS;
constexpr int kFinChunk = 4096;  // products staged per pass (32 KB of shared memory)
__global__ void __launch_bounds__(kFinThreads) finish_kernel(const double* __restrict__ g, double* blk, double* __restrict__ xc,
                                                             double2* __restrict__ pend, Scal* sc, unsigned* arrivals, long long n,
                                                             CalcConsts cc, int which, Cut cut, int no_defer, int reduce_mode,
                                                             int nranks) {
    ELL_STATIC_SHARED double prod[kFinChunk];
    ELL_STATIC_SHARED double s_omega;
    ELL_STATIC_SHARED double s_kappa_in;
    ELL_STATIC_SHARED unsigned long long s_epoch;
    ELL_STATIC_SHARED int s_timeout;
    ELL_STATIC_SHARED Scal s_out;
    ELL_STATIC_SHARED double s_roo;
    const int tid = threadIdx.x;
    pdl_wait();      // the sweep of this update is complete (its M stores too)
    pdl_trigger();   // the next update's sweep may start its prologue: barriers and the first tiles of M
    if (tid == 0) {
        s_epoch = sc->epoch;
        s_kappa_in = sc->kappa;
        s_timeout = *reinterpret_cast<volatile int*>(&sc->failed);   // sticky: a handle that lost a peer stays failed
    }
    __syncthreads();
    const unsigned long long epoch = s_epoch;
    const bool ll = nranks > 0;
    const uint4* lines = ll ? xblk_lines(blk, n, (int)(epoch & 1ULL)) : nullptr;
    const unsigned flag = (unsigned)epoch;
    const bool dead = s_timeout != 0;
    bool lost = false;   // this thread gave up on a line
    auto load_gt = [&](long long k) -> double {
        if (!ll) return __ldcg(blk + k);
        double v = 0.0;
        if (dead || lost) return v;
        if (!ll_load(lines + k, flag, v)) lost = true;
        return v;
    };
    if (reduce_mode == 0) {
        // grad.dot(&grad_t): products in parallel, the sum strictly left to right (src/arr.rs:443-451)
        double acc = 0.0;
        for (long long base = 0; base < n; base += kFinChunk) {
            const int len = (int)((n - base) < kFinChunk ? (n - base) : kFinChunk);
            for (int k = tid; k < len; k += kFinThreads) prod[k] = g[base + k] * load_gt(base + k);
            __syncthreads();
            if (tid == 0) {
                int k = 0;
                for (; k + 8 <= len; k += 8) {
                    const double p0 = prod[k], p1 = prod[k + 1], p2 = prod[k + 2], p3 = prod[k + 3];
                    const double p4 = prod[k + 4], p5 = prod[k + 5], p6 = prod[k + 6], p7 = prod[k + 7];
                    acc += p0; acc += p1; acc += p2; acc += p3; acc += p4; acc += p5; acc += p6; acc += p7;
                }
                for (; k < len; ++k) acc += prod[k];
            }
            __syncthreads();
        }
        if (tid == 0) s_omega = acc;
    } else {
        // blocked: thread t sums elements t, t+1024, … in order, then a fixed binary tree — deterministic, the same on every rank.
        // The lines are requested 8 at a time before any of them is checked: one L2 round trip per batch instead of one per line.
        double acc = 0.0;
        long long k = tid;
        if (ll && !dead) {
            for (; k + 7 * kFinThreads < n; k += 8 * kFinThreads) {
                unsigned w[8][4];
                double gk[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint4* line = lines + k + (long long)u * kFinThreads;
                    ll_peek(line, w[u]);
                    gk[u] = g[k + (long long)u * kFinThreads];
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    double v = __hiloint2double((int)w[u][2], (int)w[u][0]);
                    if (w[u][1] != flag || w[u][3] != flag) v = load_gt(k + (long long)u * kFinThreads);   // not there yet: poll
                    acc += gk[u] * v;
                }
            }
        }
#pragma unroll 4
        for (; k < n; k += kFinThreads) acc += g[k] * load_gt(k);
        // fixed tree: xor-shuffles inside each warp, the 32 warp sums through shared memory, xor-shuffles again (two barriers
        // instead of ten; the same association on every rank)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if ((tid & 31) == 0) prod[tid >> 5] = acc;
        __syncthreads();
        if (tid < 32) {
            double v = tid < kFinThreads / 32 ? prod[tid] : 0.0;   // (all 32 entries exist with the device's 1024 threads)
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (tid == 0) s_omega = v;
        }
    }
    if (lost) {
        s_timeout = 1;
        atomicExch(&sc->failed, 1);   // one decision for the whole handle: the CTA that publishes re-reads it
    }
    __syncthreads();
    if (tid == 0) {
        Scal o;
        o.epoch = nranks > 0 ? epoch + 1 : epoch;
        o.kappa = s_kappa_in;
        o.omega = 0.0; o.tsq = 0.0; o.ratio = 0.0; o.scale = 1.0; o.pending = 0; o.status = kUnknown; o.failed = 0; o._pad = 0;
        s_roo = 0.0;
        if (!s_timeout) {
            const double omega = s_omega;
            const double tsq = s_kappa_in * omega;
            o.tsq = tsq;
            o.omega = omega;
            Rsd r;
            o.status = ell_calc(cc, which, cut, tsq, r);
            if (o.status == kSuccess) {
                s_roo = r.rho / omega;
                o.ratio = r.sigma / omega;
                double kappa = s_kappa_in * r.delta;
                if (no_defer) {
                    o.scale = kappa;
                    kappa = 1.0;
                }
                o.kappa = kappa;
                o.pending = 1;
            }   // else: state unchanged except tsq (src/ell.rs:107-109)
        }
        s_out = o;
    }
    __syncthreads();
    if (s_out.status == kSuccess) {
        const double roo = s_roo, ratio = s_out.ratio;
        const long long per = (n + gridDim.x - 1) / gridDim.x;
        const long long lo = (long long)blockIdx.x * per;
        const long long hi = lo + per < n ? lo + per : n;
#pragma unroll 4
        for (long long i = lo + tid; i < hi; i += kFinThreads) {
            const double t = load_gt(i);                 // every line has arrived: no waiting here
            xc[i] -= roo * t;                            // src/ell.rs:111-115
            pend[i] = make_double2(t, ratio * t);        // r_qg = ratio * grad_t[i]  (src/ell.rs:119), applied by the next sweep
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(arrivals, 1u);
        if (prev == gridDim.x - 1) {   // last CTA: everyone has read sc
            *arrivals = 0;
            const double old_tsq = sc->tsq;
            const int failed = atomicOr(&sc->failed, 0);   // some CTA (maybe not this one) gave up waiting for a peer
            Scal o = s_out;
            if (failed) {   // the slices of xc/pend other CTAs may have written are void: the handle is dead from here on
                o.status = kUnknown; o.pending = 0; o.ratio = 0.0; o.scale = 1.0; o.kappa = s_kappa_in; o.omega = 0.0;
                o.tsq = old_tsq;
            }
            o.failed = failed;
            o._pad = 0;
            *sc = o;
        }
    }
}

__global__ void transpose_kernel(const double* __restrict__ in, long long ld_in, double* __restrict__ out, long long rows_in,
                                 long long cols_in) {
    __shared__ double t[32][33];
    const long long bx = (long long)blockIdx.x * 32, by = (long long)blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const long long i = by + r, j = bx + threadIdx.x;
        if (i < rows_in && j < cols_in) t[r][threadIdx.x] = in[i * ld_in + j];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const long long j = bx + r, i = by + threadIdx.x;  // out[j][i] = in[i][j]
        if (i < rows_in && j < cols_in) out[j * rows_in + i] = t[threadIdx.x][r];
    }
}

__global__ void fill_diag_kernel(double* __restrict__ s, long long ld, long long n, long long w, long long c_glob0,
                                 const double* __restrict__ diag) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= w) return;
    const long long i = c_glob0 + c;
    if (i < n) s[i * ld + c] = diag ? diag[i] : 1.0;
}

constexpr size_t sweep_smem(int cols, int rows, int stages) { return (size_t)stages * (cols * rows * 8 + rows * 24) + stages * 8 + 128; }

}  // namespace
}  // namespace ellcuda

namespace ellcuda {
namespace {

#ifdef __CUDACC__   // host side: left out when tests/cpp/emu_stream.cpp compiles the kernels of this file for the host
void free_ctx(ellcuda_ctx* h) {
    if (!h) return;
    for (size_t r = h->parts.size(); r-- > 0;) free_ctx(h->parts[r]);   // reverse: a later slab may borrow an earlier one's stream
    h->parts.clear();
    DevGuard guard(h->device);
    cudaFree(h->A); cudaFree(h->ytmp);
    sym_free(h);
    cudaFree(h->st_qu); cudaFree(h->st_scal); cudaFree(h->st_status);
    cudaFree(h->S); cudaFree(h->xc); cudaFree(h->gpad); cudaFree(h->gzero); cudaFree(h->gstage); cudaFreeHost(h->status_host);
    cudaFree(h->pend); cudaFree(h->partial); cudaFree(h->counters); cudaFree(h->sc);
    for (int r = 0; r < kMaxRanks; ++r)
        if (h->peer_opened[r] && h->peer_blk[r]) cudaIpcCloseMemHandle(h->peer_blk[r]);
    cudaFree(h->xblk);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int alloc_ctx(ellcuda_ctx* h) {
    const long long n = h->n;
    if (const char* e = getenv("ELLCUDA_SWEEP")) {   // performance experiments: "rows,stages"
        int r = 0, st = 0;
        if (sscanf(e, "%d,%d", &r, &st) == 2 && (r == 16 || r == 32 || r == 48 || r == 64) && (st == 3 || st == 4 || st == 6 || st == 8)) {
            h->tile_rows = r;
            h->stages = st;
        }
    }
    h->ld = (h->w + 1) & ~1LL;  // TMA needs a 16-byte row pitch
    h->npad = (n + kRowsMax - 1) / kRowsMax * kRowsMax + kRowsMax;
    // column-block width: the widest of 64 / 32 / 16 that still gives ~192 sweep warps (all resident, 2-3 per SM)
    h->cols = kColsMax;
    while (h->cols > 16 && ((h->w + h->cols - 1) / h->cols) * h->nseg < 192) h->cols /= 2;
    if (const char* e = getenv("ELLCUDA_COLS")) {   // performance experiments
        const int c = atoi(e);
        if (c == 16 || c == 32 || c == 64) h->cols = c;
    }
    if (h->cols != kColsMax) { h->tile_rows = 64; h->stages = h->cols == 32 ? 4 : 6; }
    h->ld_part = (h->w + kColsMax - 1) / kColsMax * kColsMax;
    h->ncb = (h->w + h->cols - 1) / h->cols;
    const size_t pend_n = (size_t)(h->npad + kColsMax + 2);
    if (!h->sym) ELL_CUDA_TRY(cudaMalloc(&h->S, (size_t)n * h->ld * 8));
    ELL_CUDA_TRY(cudaMalloc(&h->xc, (size_t)n * 8));
    ELL_CUDA_TRY(cudaMalloc(&h->gpad, (size_t)h->npad * 8));
    ELL_CUDA_TRY(cudaMalloc(&h->gzero, (size_t)h->npad * 8));
    ELL_CUDA_TRY(cudaMalloc(&h->gstage, (size_t)h->npad * 8));
    ELL_CUDA_TRY(cudaMallocHost(&h->status_host, sizeof(int)));
    const size_t xblk_total = xblk_doubles(n) + (h->sym && h->sym_count > 1 ? (size_t)4 * h->sym_count * ((n + h->sym_count - 1) / h->sym_count) : 0);
    ELL_CUDA_TRY(cudaMalloc(&h->xblk, xblk_total * 8));
    ELL_CUDA_TRY(cudaMalloc(&h->pend, pend_n * 16));
    ELL_CUDA_TRY(cudaMalloc(&h->partial, (size_t)h->nseg * h->ld_part * 8));
    ELL_CUDA_TRY(cudaMalloc(&h->counters, (size_t)(h->ncb + 2) * 4));
    ELL_CUDA_TRY(cudaMalloc(&h->sc, sizeof(Scal)));
    ELL_CUDA_TRY(cudaEventCreate(&h->ev0));
    ELL_CUDA_TRY(cudaEventCreate(&h->ev1));
    ELL_CUDA_TRY(cudaMemset(h->gpad, 0, (size_t)h->npad * 8));
    ELL_CUDA_TRY(cudaMemset(h->gzero, 0, (size_t)h->npad * 8));
    ELL_CUDA_TRY(cudaMemset(h->gstage, 0, (size_t)h->npad * 8));
    ELL_CUDA_TRY(cudaMemset(h->xblk, 0, xblk_total * 8));
    ELL_CUDA_TRY(cudaMemset(h->pend, 0, pend_n * 16));
    ELL_CUDA_TRY(cudaMemset(h->counters, 0, (size_t)(h->ncb + 2) * 4));
    keep_mempool(h->device);
    if (h->sym) return sym_alloc(h);
    return make_tensor_map(&h->map, h->S, n, h->w, h->ld, h->tile_rows, h->cols);
}

template <int CO, int R, int ST, int W>
int launch_sweep_t(ellcuda_ctx* h, dim3 grid, const SweepArgs& a, cudaStream_t s) {
    // the opt-in to > 48 KB of dynamic shared memory is a per-DEVICE function attribute: one bit per device ordinal,
    // per instantiation, set under a mutex (handles on several devices / host threads in one process)
    static std::mutex mu;
    static unsigned long long done_mask = 0;
    constexpr size_t smem = sweep_smem(CO, R, ST);
    {
        std::lock_guard<std::mutex> lk(mu);
        const bool known = h->device < 64 && ((done_mask >> h->device) & 1ULL);
        if (!known) {
            ELL_CUDA_TRY(cudaFuncSetAttribute(sweep_kernel<CO, R, ST, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (h->device < 64) done_mask |= 1ULL << h->device;
        }
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = dim3(32 * W); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: the prologue (barriers, first tiles of M) may
    attr[0].val.programmaticStreamSerializationAllowed = 1;            // overlap the finish kernel of the previous update
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    ELL_CUDA_TRY(cudaLaunchKernelEx(&cfg, sweep_kernel<CO, R, ST, W>, h->asym_live ? h->mapA : h->map, a));
    return ELLCUDA_OK;
}

int launch_sweep(ellcuda_ctx* h, dim3 grid, const SweepArgs& a, cudaStream_t s) {
    if (h->cols == 32) return launch_sweep_t<32, 64, 4, 1>(h, grid, a, s);
    if (h->cols == 16) return launch_sweep_t<16, 64, 6, 1>(h, grid, a, s);
    const int key = h->tile_rows * 100 + h->stages;   // 64-column blocks: tile rows x ring depth (ELLCUDA_SWEEP, experiments)
    switch (key) {
        case 1608: return launch_sweep_t<64, 16, 8, 1>(h, grid, a, s);
        case 3204: return launch_sweep_t<64, 32, 4, 1>(h, grid, a, s);
        case 3206: return launch_sweep_t<64, 32, 6, 1>(h, grid, a, s);
        case 4804: return launch_sweep_t<64, 48, 4, 1>(h, grid, a, s);
        default:
            // one warp per CTA with two columns per lane; ELLCUDA_SWEEP_WARPS=2 selects two warps with one column per lane
            // (measured at n=16384: 711 us per sweep against 680 on the whole matrix, 100.4 against 99.5 us on a 1/8 slab)
            if (ELL_ENV_INT("ELLCUDA_SWEEP_WARPS", 1) == 2) return launch_sweep_t<64, 64, 3, 2>(h, grid, a, s);
            return launch_sweep_t<64, 64, 3, 1>(h, grid, a, s);
    }
}

// begin = [padded copy of g] + sweep (deferred rank-1, column sums, ordered combine, push to peers).
// g_dev may be null (flush: applies the pending rank-1 only, produces no grad_t).
int enqueue_begin(ellcuda_ctx* h, const double* g_dev, double* plain_out = nullptr) {
    cudaStream_t s = h->stream;
    const bool flush = g_dev == nullptr;
    if (flush) {
        h->gsrc = h->gzero;
    } else if (h->n % kRowsMax == 0 && h->n % h->tile_rows == 0 && (reinterpret_cast<uintptr_t>(g_dev) & 15) == 0) {
        h->gsrc = g_dev;      // tiles never read past n: the caller's buffer is loaded directly (no copy kernel)
    } else {
        const int thr = 256;
        copy_pad_kernel<<<(unsigned)((h->npad + thr - 1) / thr), thr, 0, s>>>(g_dev, h->gpad, h->n, h->npad);
        h->gsrc = h->gpad;
        h->launches += 1;
    }
    if (h->sym) {   // symmetric storage: block-triangle sweep + ordered combine (stream_sym.cu)
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(s, &cap);
        const unsigned evflag = cap == cudaStreamCaptureStatusNone ? cudaEventRecordDefault : cudaEventRecordExternal;
        if (h->timing) ELL_CUDA_TRY(cudaEventRecordWithFlags(h->ev0, s, evflag));
        const int rc = sym_sweep(h, h->gsrc, flush ? 0 : 1, plain_out);
        if (rc) return rc;
        if (h->timing) { ELL_CUDA_TRY(cudaEventRecordWithFlags(h->ev1, s, evflag)); h->timed = true; }
        return ELLCUDA_OK;
    }
    SweepArgs a;
    a.g = h->gsrc; a.pend = h->pend; a.sc = h->sc; a.partial = h->partial;
    a.cnt_block = h->counters; a.own_blk = h->xblk; a.out = plain_out;
    a.ld_part = h->ld_part; a.n = h->n; a.w = h->w; a.c_glob0 = h->c0; a.nseg = h->nseg;
    a.exchange = flush ? 0 : 1;
    a.pt.nranks = h->nranks; a.pt.rank = h->rank;
    for (int r = 0; r < kMaxRanks; ++r) a.pt.blk[r] = r < h->nranks ? h->peer_blk[r] : nullptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &cap);
    // inside a stream capture the records become external event-record nodes, so a replayed graph still times its sweeps
    const unsigned evflag = cap == cudaStreamCaptureStatusNone ? cudaEventRecordDefault : cudaEventRecordExternal;
    if (h->timing) ELL_CUDA_TRY(cudaEventRecordWithFlags(h->ev0, s, evflag));
    dim3 grid((unsigned)h->ncb, (unsigned)h->nseg);
    int rc = launch_sweep(h, grid, a, s);
    if (rc) return rc;
    if (h->timing) { ELL_CUDA_TRY(cudaEventRecordWithFlags(h->ev1, s, evflag)); h->timed = true; }
    ELL_CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return ELLCUDA_OK;
}
#endif  // __CUDACC__


__global__ void clear_pending_kernel(Scal* sc) { sc->pending = 0; sc->scale = 1.0; }

// sum of v[0..n) strictly left to right from 0.0 (Arr::dot / Arr::sum order, src/arr.rs:443-451): the chunks are staged in
// shared memory by the whole CTA, one thread walks the chain (10 clocks per element: the FP64 add latency)
__global__ void __launch_bounds__(1024) seq_sum_kernel(const double* __restrict__ v, long long n, double* __restrict__ out) {
    __shared__ double buf[4096];
    double acc = 0.0;
    for (long long base = 0; base < n; base += 4096) {
        const int len = (int)((n - base) < 4096 ? (n - base) : 4096);
        for (int k = threadIdx.x; k < len; k += 1024) buf[k] = v[base + k];
        __syncthreads();
        if (threadIdx.x == 0)
            for (int k = 0; k < len; ++k) acc += buf[k];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = acc;
}

#ifdef __CUDACC__

int flush_pending(ellcuda_ctx* h) {  // make S hold the reference's `mq` exactly
    int rc = enqueue_begin(h, nullptr);
    if (rc) return rc;
    clear_pending_kernel<<<1, 1, 0, h->stream>>>(h->sc);
    ELL_CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return ELLCUDA_OK;
}

#endif  // __CUDACC__

}  // namespace
}  // namespace ellcuda

using namespace ellcuda;

#ifdef __CUDACC__
extern "C" {
// ---- EllStable as a single host-driven search space (src/ell_stable.rs:9-35, 128-210), n <= 2048 -----------------------
static int create_stable(const ell_create_desc* d, ell_handle* out) {
    if (d->n > 2048) { set_error("a single EllStable handle supports n <= 2048 (its recurrences are serial; the throughput path is batched mode)"); return ELLCUDA_E_ARG; }
    if ((d->row_begin != 0 || (d->row_end != 0 && d->row_end != d->n))) { set_error("EllStable does not shard"); return ELLCUDA_E_ARG; }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (libellcuda has no CPU fallback)");
        return ELLCUDA_E_CUDA;
    }
    if (d->device < 0 || d->device >= count) { set_error("device ordinal out of range"); return ELLCUDA_E_ARG; }
    cudaDeviceProp prop;
    ELL_CUDA_TRY(cudaGetDeviceProperties(&prop, d->device));
    if (prop.major < 10) { set_error("libellcuda is built for sm_100a (B200) only"); return ELLCUDA_E_CUDA; }
    ELL_DEVICE_SCOPE(d->device);
    const long long n = d->n, tri = n * (n + 1) / 2;
    ellcuda_ctx* h = new ellcuda_ctx();
    h->kind = ELL_KIND_ELLSTABLE;
    h->device = d->device; h->n = n; h->w = n; h->use_parallel = d->use_parallel_cut;
    if (d->stream) h->stream = (cudaStream_t)d->stream;
    else {
        cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
        h->own_stream = true;
    }
    std::vector<double> qu((size_t)tri, 0.0);   // strict upper = L-factor entries, diagonal = D (src/ell_stable.rs:52-125)
    for (long long r = 0, e = 0; r < n; ++r)
        for (long long c = r; c < n; ++c, ++e)
            qu[e] = d->mq ? d->mq[r * n + c] : (r == c ? (d->diag ? d->diag[r] : 1.0) : 0.0);
    const double scal[2] = {d->kappa, 0.0};
    cudaError_t e = cudaMalloc(&h->st_qu, (size_t)tri * 8);
    if (e == cudaSuccess) e = cudaMalloc(&h->st_scal, 16);
    if (e == cudaSuccess) e = cudaMalloc(&h->st_status, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&h->xc, (size_t)n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&h->gstage, (size_t)n * 8);
    if (e == cudaSuccess) e = cudaMallocHost(&h->status_host, sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->st_qu, qu.data(), (size_t)tri * 8, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->st_scal, scal, 16, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(h->st_status, 0, sizeof(int), h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->xc, d->xc, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { free_ctx(h); return cuda_fail(e, "ellcuda_create (EllStable)", __FILE__, __LINE__); }
    *out = h;
    return ELLCUDA_OK;
}

int ellcuda_create(const ell_create_desc* d, ell_handle* out) {
    if (!d || !out) { set_error("null argument"); return ELLCUDA_E_ARG; }
    *out = nullptr;
    if (d->kind != ELL_KIND_ELL && d->kind != ELL_KIND_ELLSTABLE) { set_error("bad kind"); return ELLCUDA_E_ARG; }
    if (d->n <= 0 || !d->xc) { set_error("n must be positive and xc given"); return ELLCUDA_E_ARG; }
    if (d->kind == ELL_KIND_ELLSTABLE) return create_stable(d, out);
    long long rb = d->row_begin, re = d->row_end;
    if (rb == 0 && re == 0) re = d->n;
    if (rb < 0 || re > d->n || rb >= re) { set_error("bad row range"); return ELLCUDA_E_ARG; }
    if (d->reduce_mode < 0 || d->reduce_mode > 64) { set_error("reduce_mode must be in [0, 64]"); return ELLCUDA_E_ARG; }
    if (d->storage != ELL_STORAGE_FULL && d->storage != ELL_STORAGE_SYMMETRIC) { set_error("bad storage"); return ELLCUDA_E_ARG; }
    const bool sym = d->storage == ELL_STORAGE_SYMMETRIC;
    if (sym && (rb != 0 || re != d->n)) { set_error("symmetric storage shards by shard_rank / shard_count, not by rows"); return ELLCUDA_E_ARG; }
    const int shard_count = d->shard_count > 1 ? d->shard_count : 1;
    if (shard_count > 1 && (!sym || shard_count > kMaxRanks || d->shard_rank < 0 || d->shard_rank >= shard_count)) {
        set_error("shard_rank / shard_count apply to ELL_STORAGE_SYMMETRIC, 0 <= rank < count <= 16");
        return ELLCUDA_E_ARG;
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (libellcuda has no CPU fallback)");
        return ELLCUDA_E_CUDA;
    }
    if (d->device < 0 || d->device >= count) { set_error("device ordinal out of range"); return ELLCUDA_E_ARG; }
    cudaDeviceProp prop;
    ELL_CUDA_TRY(cudaGetDeviceProperties(&prop, d->device));
    if (prop.major < 10) { set_error("libellcuda is built for sm_100a (B200) only"); return ELLCUDA_E_CUDA; }
    ELL_DEVICE_SCOPE(d->device);
    bool asym = false;
    if (d->mq) {  // Ell::new_with_matrix takes any Arr (src/ell.rs:31-41); a non-symmetric one needs the `A` copy (see ellcuda_ctx)
        const long long n = d->n;
        for (long long i = 0; i < n && !asym; ++i)
            for (long long j = 0; j < i; ++j)
                if (!(d->mq[i * n + j] == d->mq[j * n + i]) && !(d->mq[i * n + j] != d->mq[i * n + j] && d->mq[j * n + i] != d->mq[j * n + i])) {
                    asym = true;
                    break;
                }
    }
    if (sym && asym) { set_error("symmetric storage needs a symmetric matrix"); return ELLCUDA_E_ARG; }
    ellcuda_ctx* h = new ellcuda_ctx();
    h->device = d->device;
    h->sym = sym;
    h->sym_rank = shard_count > 1 ? d->shard_rank : 0;
    h->sym_count = shard_count;
    h->n = d->n; h->c0 = rb; h->w = re - rb;
    h->reduce_mode = d->reduce_mode;
    h->nseg = d->reduce_mode > 0 ? d->reduce_mode : 1;
    h->no_defer = d->no_defer_trick; h->use_parallel = d->use_parallel_cut;
    if (d->stream) h->stream = (cudaStream_t)d->stream;
    else {
        cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
        h->own_stream = true;
    }
    int rc = alloc_ctx(h);
    if (rc) { free_ctx(h); return rc; }
    cudaStream_t s = h->stream;
    cudaError_t e = cudaSuccess;
    const long long n = h->n;
    std::vector<double> stage_l, stage_a;   // must outlive the asynchronous copies below
    if (sym) {
        rc = sym_upload(h, d->mq, d->diag);
        if (rc) { free_ctx(h); return rc; }
    } else if (d->mq && asym) {
        // S <- column block of the mirrored lower triangle; A <- the slab's rows as given, transposed
        const long long w = h->w, c0 = h->c0;
        stage_l.resize((size_t)n * w);
        stage_a.resize((size_t)n * w);
        for (long long i = 0; i < n; ++i)
            for (long long c = 0; c < w; ++c) {
                const long long j = c0 + c;
                stage_l[(size_t)i * w + c] = i >= j ? d->mq[i * n + j] : d->mq[j * n + i];
                stage_a[(size_t)i * w + c] = d->mq[j * n + i];
            }
        e = cudaMalloc(&h->A, (size_t)n * h->ld * 8);
        if (e == cudaSuccess) e = cudaMemsetAsync(h->A, 0, (size_t)n * h->ld * 8, s);
        if (e == cudaSuccess) e = cudaMemsetAsync(h->S, 0, (size_t)n * h->ld * 8, s);
        if (e == cudaSuccess) e = cudaMemcpy2DAsync(h->S, (size_t)h->ld * 8, stage_l.data(), (size_t)w * 8, (size_t)w * 8, (size_t)n, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpy2DAsync(h->A, (size_t)h->ld * 8, stage_a.data(), (size_t)w * 8, (size_t)w * 8, (size_t)n, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) {
            rc = make_tensor_map(&h->mapA, h->A, n, h->w, h->ld, h->tile_rows, h->cols);
            if (rc) { cudaStreamSynchronize(s); free_ctx(h); return rc; }
            h->asym_live = true;
        }
    } else if (d->mq) {  // column block [c0, c0+w) of the row-major host matrix
        e = cudaMemcpy2DAsync(h->S, (size_t)h->ld * 8, d->mq + h->c0, (size_t)n * 8, (size_t)h->w * 8, (size_t)n, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess && h->ld != h->w) e = cudaMemset2DAsync(h->S + h->w, (size_t)h->ld * 8, 0, 8, (size_t)n, s);
    } else {
        e = cudaMemsetAsync(h->S, 0, (size_t)n * h->ld * 8, s);
        double* ddiag = nullptr;
        if (e == cudaSuccess && d->diag) {
            e = cudaMallocAsync((void**)&ddiag, (size_t)n * 8, s);
            if (e == cudaSuccess) e = cudaMemcpyAsync(ddiag, d->diag, (size_t)n * 8, cudaMemcpyHostToDevice, s);
        }
        if (e == cudaSuccess) {
            fill_diag_kernel<<<(unsigned)((h->w + 255) / 256), 256, 0, s>>>(h->S, h->ld, n, h->w, h->c0, ddiag);
            e = cudaGetLastError();
        }
        if (ddiag) cudaFreeAsync(ddiag, s);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->xc, d->xc, (size_t)n * 8, cudaMemcpyHostToDevice, s);
    Scal sc0;
    memset(&sc0, 0, sizeof sc0);
    sc0.kappa = d->kappa; sc0.scale = 1.0; sc0.epoch = 1;
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->sc, &sc0, sizeof sc0, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { free_ctx(h); return cuda_fail(e, "ellcuda_create upload", __FILE__, __LINE__); }
    *out = h;
    return ELLCUDA_OK;
}

static int attach_group(ellcuda_ctx* g);

int ellcuda_clone(ell_handle h, ell_handle* out) {
    if (!h || !out) { set_error("null handle"); return ELLCUDA_E_ARG; }
    if (!h->parts.empty()) {   // group: clone every slab, then attach the clones to each other
        ellcuda_ctx* g = new ellcuda_ctx();
        g->kind = h->kind; g->device = h->device; g->n = h->n; g->w = h->w;
        g->reduce_mode = h->reduce_mode; g->no_defer = h->no_defer; g->use_parallel = h->use_parallel;
        for (ellcuda_ctx* p : h->parts) {
            ell_handle c = nullptr;
            int rc = ellcuda_clone(p, &c);
            if (rc) { free_ctx(g); return rc; }
            for (ellcuda_ctx* q : g->parts)   // slabs on one device share one stream (see ellcuda_create_sharded)
                if (q->device == c->device) { rc = ellcuda_set_stream(c, q->stream); break; }
            g->parts.push_back(c);
            if (rc) { free_ctx(g); return rc; }
        }
        const int rc = attach_group(g);
        if (rc) { free_ctx(g); return rc; }
        *out = g;
        return ELLCUDA_OK;
    }
    ELL_DEVICE_SCOPE(h->device);
    if (h->kind == ELL_KIND_ELLSTABLE) {
        const long long n = h->n, tri = n * (n + 1) / 2;
        ellcuda_ctx* c = new ellcuda_ctx();
        c->kind = h->kind; c->device = h->device; c->n = n; c->w = n; c->use_parallel = h->use_parallel;
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
        c->own_stream = true;
        e = cudaMalloc(&c->st_qu, (size_t)tri * 8);
        if (e == cudaSuccess) e = cudaMalloc(&c->st_scal, 16);
        if (e == cudaSuccess) e = cudaMalloc(&c->st_status, sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&c->xc, (size_t)n * 8);
        if (e == cudaSuccess) e = cudaMalloc(&c->gstage, (size_t)n * 8);
        if (e == cudaSuccess) e = cudaMallocHost(&c->status_host, sizeof(int));
        cudaStream_t s = h->stream;
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->st_qu, h->st_qu, (size_t)tri * 8, cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->st_scal, h->st_scal, 16, cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->st_status, h->st_status, sizeof(int), cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->xc, h->xc, (size_t)n * 8, cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { free_ctx(c); return cuda_fail(e, "ellcuda_clone (EllStable)", __FILE__, __LINE__); }
        *out = c;
        return ELLCUDA_OK;
    }
    ellcuda_ctx* c = new ellcuda_ctx();
    c->device = h->device; c->n = h->n; c->w = h->w; c->c0 = h->c0;
    c->nseg = h->nseg; c->reduce_mode = h->reduce_mode; c->no_defer = h->no_defer; c->use_parallel = h->use_parallel;
    c->sym = h->sym; c->sym_rank = h->sym_rank; c->sym_count = h->sym_count;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
    c->own_stream = true;
    int rc = alloc_ctx(c);
    if (rc) { free_ctx(c); return rc; }
    cudaStream_t s = h->stream;
    const long long n = h->n;
    // the clone is detached from peers; the deferred rank-1 travels as (pend, sc.pending/ratio/scale)
    Scal sc_h;
    if (h->sym) e = sym_copy(c, h, s) == ELLCUDA_OK ? cudaSuccess : cudaErrorUnknown;
    else e = cudaMemcpyAsync(c->S, h->S, (size_t)n * h->ld * 8, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->xc, h->xc, (size_t)n * 8, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->pend, h->pend, (size_t)(h->npad + kColsMax + 2) * 16, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess && h->asym_live) {
        e = cudaMalloc(&c->A, (size_t)n * h->ld * 8);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->A, h->A, (size_t)n * h->ld * 8, cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess) {
            rc = make_tensor_map(&c->mapA, c->A, n, c->w, c->ld, c->tile_rows, c->cols);
            if (rc) { cudaStreamSynchronize(s); free_ctx(c); return rc; }
            c->asym_live = true;
        }
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&sc_h, h->sc, sizeof(Scal), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    sc_h.epoch = 1;
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->sc, &sc_h, sizeof(Scal), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { free_ctx(c); return cuda_fail(e, "ellcuda_clone copy", __FILE__, __LINE__); }
    *out = c;
    return ELLCUDA_OK;
}

void ellcuda_destroy(ell_handle h) { free_ctx(h); }

int ellcuda_xc(ell_handle h, double* out) {
    if (!h || !out) { set_error("null argument"); return ELLCUDA_E_ARG; }
    if (!h->parts.empty()) return ellcuda_xc(h->parts[0], out);   // xc is replicated, identical on every slab handle
    ELL_DEVICE_SCOPE(h->device);
    ELL_CUDA_TRY(cudaMemcpyAsync(out, h->xc, (size_t)h->n * 8, cudaMemcpyDeviceToHost, h->stream));
    ELL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return ELLCUDA_OK;
}

int ellcuda_set_xc(ell_handle h, const double* x) {
    if (!h || !x) { set_error("null argument"); return ELLCUDA_E_ARG; }
    if (!h->parts.empty()) {
        for (ellcuda_ctx* p : h->parts) { const int rc = ellcuda_set_xc(p, x); if (rc) return rc; }
        return ELLCUDA_OK;
    }
    ELL_DEVICE_SCOPE(h->device);
    ELL_CUDA_TRY(cudaMemcpyAsync(h->xc, x, (size_t)h->n * 8, cudaMemcpyHostToDevice, h->stream));
    ELL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return ELLCUDA_OK;
}

static int read_scal(ell_handle h, Scal* sc) {
    ELL_DEVICE_SCOPE(h->device);
    ELL_CUDA_TRY(cudaMemcpyAsync(sc, h->sc, sizeof(Scal), cudaMemcpyDeviceToHost, h->stream));
    ELL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return ELLCUDA_OK;
}

static int read_stable_scal(ell_handle h, int which, double* out) {
    ELL_DEVICE_SCOPE(h->device);
    double v[2];
    ELL_CUDA_TRY(cudaMemcpyAsync(v, h->st_scal, 16, cudaMemcpyDeviceToHost, h->stream));
    ELL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    *out = v[which];
    return ELLCUDA_OK;
}

int ellcuda_tsq(ell_handle h, double* out) {
    if (!h || !out) { set_error("null argument"); return ELLCUDA_E_ARG; }
    if (!h->parts.empty()) return ellcuda_tsq(h->parts[0], out);
    if (h->kind == ELL_KIND_ELLSTABLE) return read_stable_scal(h, 1, out);
    Scal sc;
    int rc = read_scal(h, &sc);
    if (rc) return rc;
    *out = sc.tsq;
    return ELLCUDA_OK;
}

int ellcuda_kappa(ell_handle h, double* out) {
    if (!h || !out) { set_error("null argument"); return ELLCUDA_E_ARG; }
    if (!h->parts.empty()) return ellcuda_kappa(h->parts[0], out);
    if (h->kind == ELL_KIND_ELLSTABLE) return read_stable_scal(h, 0, out);
    Scal sc;
    int rc = read_scal(h, &sc);
    if (rc) return rc;
    *out = sc.kappa;
    return ELLCUDA_OK;
}

int ellcuda_get_mq(ell_handle h, double* out) {  // this slab's rows, w × n row-major
    if (!h || !out) { set_error("null argument"); return ELLCUDA_E_ARG; }
    if (!h->parts.empty() && h->parts[0]->sym) {   // symmetric group: every 64x64 tile from the slab that owns its column-block pair
        const long long n = h->n, T = (n + 63) / 64;
        const int P = (int)h->parts.size();
        std::vector<double> tmp((size_t)n * n);
        for (int r = 0; r < P; ++r) {
            const int rc = ellcuda_get_mq(h->parts[r], tmp.data());
            if (rc) return rc;
            for (long long i = 0; i < n; ++i)
                for (long long j = 0; j < n; ++j) {
                    const long long J = (i < j ? i : j) / 64;            // column block of the stored (lower) tile
                    const long long pair = J < T - 1 - J ? J : T - 1 - J;
                    if (pair % P == r) out[i * n + j] = tmp[(size_t)i * n + j];
                }
        }
        return ELLCUDA_OK;
    }
    if (!h->parts.empty()) {   // group: the slabs in row order make up the whole matrix
        for (ellcuda_ctx* p : h->parts) { const int rc = ellcuda_get_mq(p, out + p->c0 * p->n); if (rc) return rc; }
        return ELLCUDA_OK;
    }
    ELL_DEVICE_SCOPE(h->device);
    if (h->kind == ELL_KIND_ELLSTABLE) {   // the factor: diagonal + strict upper; the strict lower (scratch in the reference) is 0
        const long long n = h->n, tri = n * (n + 1) / 2;
        std::vector<double> qu((size_t)tri);
        ELL_CUDA_TRY(cudaMemcpyAsync(qu.data(), h->st_qu, (size_t)tri * 8, cudaMemcpyDeviceToHost, h->stream));
        ELL_CUDA_TRY(cudaStreamSynchronize(h->stream));
        for (long long r = 0, e = 0; r < n; ++r)
            for (long long c = 0; c < n; ++c) out[r * n + c] = c >= r ? qu[e++] : 0.0;
        return ELLCUDA_OK;
    }
    Scal sc;
    int rc = read_scal(h, &sc);
    if (rc) return rc;
    if (sc.pending && (rc = flush_pending(h))) return rc;
    double* tmp = nullptr;
    ELL_CUDA_TRY(cudaMallocAsync((void**)&tmp, (size_t)h->w * h->n * 8, h->stream));
    cudaError_t e = cudaSuccess;
    if (h->sym) {
        if (sym_unpack(h, tmp)) e = cudaErrorUnknown;
    } else {
        dim3 grid((unsigned)((h->w + 31) / 32), (unsigned)((h->n + 31) / 32)), block(32, 8);
        transpose_kernel<<<grid, block, 0, h->stream>>>(h->asym_live ? h->A : h->S, h->ld, tmp, h->n, h->w);
        h->launches += 1;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, tmp, (size_t)h->w * h->n * 8, cudaMemcpyDeviceToHost, h->stream);
    cudaFreeAsync(tmp, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return cuda_fail(e, "ellcuda_get_mq", __FILE__, __LINE__);
    return ELLCUDA_OK;
}

// Sequential (left-to-right) sum of a device vector on the handle's device and stream; synchronous, the result comes back to the
// host.  With elementwise products formed by the caller this is Arr::dot (src/arr.rs:443-451) / the norms of
// src/power_iteration.rs:22-25 for vectors that stay on the device between two ellcuda_dot_mv calls.
int ellcuda_seq_sum(ell_handle h, const double* v_dev, int64_t n, double* out_host) {
    if (!h || !out_host || n < 0 || (n > 0 && !v_dev)) { set_error("bad arguments"); return ELLCUDA_E_ARG; }
    if (!h->parts.empty()) h = h->parts[0];
    ELL_DEVICE_SCOPE(h->device);
    if (n == 0) { *out_host = 0.0; return ELLCUDA_OK; }
    double* d = nullptr;
    ELL_CUDA_TRY(cudaMallocAsync((void**)&d, 8, h->stream));
    seq_sum_kernel<<<1, 1024, 0, h->stream>>>(v_dev, n, d);
    h->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, d, 8, cudaMemcpyDeviceToHost, h->stream);
    cudaFreeAsync(d, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return cuda_fail(e, "ellcuda_seq_sum", __FILE__, __LINE__);
    return ELLCUDA_OK;
}

int ellcuda_dot_mv(ell_handle h, const double* x_dev, double* y_dev) {
    if (!h || !x_dev || !y_dev) { set_error("null argument"); return ELLCUDA_E_ARG; }
    if (h->kind != ELL_KIND_ELL) { set_error("dot_mv needs an Ell handle"); return ELLCUDA_E_ARG; }
    if (h->sym && h->sym_count > 1) { set_error("dot_mv is not available on a sharded symmetric-storage handle"); return ELLCUDA_E_ARG; }
    if (!h->parts.empty() && h->parts[0]->sym) { set_error("dot_mv is not available on a sharded symmetric-storage handle"); return ELLCUDA_E_ARG; }
    if (!h->parts.empty()) {   // group: every device forms its rows of y from its own copy of x; y_dev receives all n entries
        for (ellcuda_ctx* p : h->parts) {
            ELL_DEVICE_SCOPE(p->device);
            if (!p->ytmp) ELL_CUDA_TRY(cudaMalloc(&p->ytmp, (size_t)p->ld_part * 8));
            ELL_CUDA_TRY(cudaMemcpyAsync(p->gstage, x_dev, (size_t)p->n * 8, cudaMemcpyDefault, p->stream));
            int rc = ellcuda_dot_mv(p, p->gstage, p->ytmp);
            if (rc) return rc;
            ELL_CUDA_TRY(cudaMemcpyAsync(y_dev + p->c0, p->ytmp, (size_t)p->w * 8, cudaMemcpyDefault, p->stream));
        }
        for (ellcuda_ctx* p : h->parts) { const int rc = ellcuda_sync(p); if (rc) return rc; }
        return ELLCUDA_OK;
    }
    ELL_DEVICE_SCOPE(h->device);
    Scal sc;
    int rc = read_scal(h, &sc);
    if (rc) return rc;
    if (sc.pending && (rc = flush_pending(h))) return rc;   // the product is with the reference's `mq`, all updates applied
    const double* keep = h->gsrc;
    rc = enqueue_begin(h, x_dev, y_dev);
    h->gsrc = keep;
    return rc;
}

int ellcuda_update_begin(ell_handle h, int32_t which, const double* grad_dev, double beta0, int32_t has_beta1, double beta1) {
    if (!h || !grad_dev) { set_error("null argument"); return ELLCUDA_E_ARG; }
    if (which < ELL_UPDATE_BIAS || which > ELL_UPDATE_Q) { set_error("bad update selector"); return ELLCUDA_E_ARG; }
    NvtxRange nvtx_range("ellcuda_update_begin (sweep)");
    if (!h->parts.empty()) {   // group: the gradient is copied to every device (it must be complete when this is called), then
        for (ellcuda_ctx* p : h->parts) {   // every slab's sweep is enqueued before any finish can start waiting for peers
            ELL_DEVICE_SCOPE(p->device);
            ELL_CUDA_TRY(cudaMemcpyAsync(p->gstage, grad_dev, (size_t)p->n * 8, cudaMemcpyDefault, p->stream));
            p->which = which;
            p->cut = has_beta1 > 0 ? parallel_cut(beta0, beta1) : single_cut(beta0);
            const int rc = enqueue_begin(p, p->gstage);
            if (rc) return rc;
        }
        for (ellcuda_ctx* p : h->parts)   // sharded symmetric storage: phase 2 of the reduction only after every slab's phase 1 is enqueued
            if (p->sym) { ELL_DEVICE_SCOPE(p->device); const int rc = sym_reduce(p); if (rc) return rc; }
        return ELLCUDA_OK;
    }
    ELL_DEVICE_SCOPE(h->device);
    h->which = which;
    h->cut = has_beta1 > 0 ? parallel_cut(beta0, beta1) : single_cut(beta0);
    if (h->kind == ELL_KIND_ELLSTABLE) {   // the whole update_core is one single-warp kernel; finish has nothing left to do
        h->launches += 1;
        return launch_stable_single(h->st_qu, h->xc, h->st_scal, h->st_status, grad_dev, (int)h->n, h->use_parallel, which, h->cut,
                                    h->stream);
    }
    int rc = enqueue_begin(h, grad_dev);
    if (rc == ELLCUDA_OK && h->sym && h->sym_count > 1) rc = sym_reduce(h);   // peers are other GPUs: their phase 1 runs concurrently
    return rc;
}

int ellcuda_update_finish(ell_handle h) {
    if (!h) { set_error("null handle"); return ELLCUDA_E_ARG; }
    NvtxRange nvtx_range("ellcuda_update_finish");
    if (!h->parts.empty()) {
        for (ellcuda_ctx* p : h->parts) { const int rc = ellcuda_update_finish(p); if (rc) return rc; }
        return ELLCUDA_OK;
    }
    if (h->kind == ELL_KIND_ELLSTABLE) return ELLCUDA_OK;
    ELL_DEVICE_SCOPE(h->device);
    if (!h->gsrc) { set_error("ellcuda_update_finish without ellcuda_update_begin"); return ELLCUDA_E_ARG; }
    const CalcConsts cc = CalcConsts::make((int)h->n, h->use_parallel);
    long long fin_ctas = (h->n + kFinThreads - 1) / kFinThreads;
    if (fin_ctas > 16) fin_ctas = 16;
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)fin_ctas); cfg.blockDim = dim3(kFinThreads); cfg.dynamicSmemBytes = 0; cfg.stream = h->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: resident while the sweep drains
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
        ELL_CUDA_TRY(cudaLaunchKernelEx(&cfg, finish_kernel, (const double*)h->gsrc, h->xblk, h->xc, h->pend, h->sc,
                                        h->counters + h->ncb + 1, h->n, cc, h->which, h->cut, h->no_defer,
                                        h->sym && h->reduce_mode == 0 ? 1 : h->reduce_mode, h->nranks));
    }
    ELL_CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    if (h->asym_live) {   // non-symmetric start: the host must learn of the first Success (see ellcuda_ctx::asym_live)
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(h->stream, &cap);
        if (cap != cudaStreamCaptureStatusNone) {
            set_error("a handle created from a non-symmetric matrix cannot be captured in a CUDA graph before its first successful cut");
            return ELLCUDA_E_ARG;
        }
        int32_t st = ELL_UNKNOWN;
        const int rc = ellcuda_last_status(h, &st);
        if (rc) return rc;
        if (st == ELL_SUCCESS) {
            h->asym_live = false;
            cudaFree(h->A);
            h->A = nullptr;
        }
    }
    return ELLCUDA_OK;
}

int ellcuda_set_no_defer_trick(ell_handle h, int32_t on) {   // `pub no_defer_trick` written after new(), src/ell.rs:10, 342-344
    if (!h) { set_error("null handle"); return ELLCUDA_E_ARG; }
    if (h->kind != ELL_KIND_ELL) { set_error("EllStable has no no_defer_trick (src/ell_stable.rs:9-15)"); return ELLCUDA_E_ARG; }
    h->no_defer = on != 0;   // read by the next finish (src/ell.rs:132-135); updates already made are not revisited, as in the reference
    for (ellcuda_ctx* p : h->parts) p->no_defer = h->no_defer;
    return ELLCUDA_OK;
}

int ellcuda_set_use_parallel_cut(ell_handle h, int32_t on) {   // EllCalc::use_parallel_cut, src/ell_calc.rs:630; Ell/EllStable helper
    if (!h) { set_error("null handle"); return ELLCUDA_E_ARG; }
    h->use_parallel = on != 0;
    for (ellcuda_ctx* p : h->parts) p->use_parallel = h->use_parallel;
    return ELLCUDA_OK;
}

int ellcuda_update(ell_handle h, int32_t which, const double* grad_host, double beta0, int32_t has_beta1, double beta1,
                   int32_t* out_status) {
    if (!h || !grad_host || !out_status) { set_error("null argument"); return ELLCUDA_E_ARG; }
    if (!h->parts.empty()) {   // group: host gradient to every device, all sweeps, all finishes, then the (identical) status
        nvtxRangePushA("ellcuda_update[group]");
        int rc = ELLCUDA_OK;
        rc = ellcuda_update_begin(h, which, grad_host, beta0, has_beta1, beta1);   // cudaMemcpyDefault: a host pointer is fine
        if (!rc) rc = ellcuda_update_finish(h);
        if (!rc) rc = ellcuda_last_status(h, out_status);
        nvtxRangePop();
        return rc;
    }
    if (h->w != h->n) { set_error("ellcuda_update needs a whole-matrix handle; sharded handles use update_begin/finish"); return ELLCUDA_E_ARG; }
    NvtxRange nvtx_range("ellcuda_update");
    ELL_DEVICE_SCOPE(h->device);
    ELL_CUDA_TRY(cudaMemcpyAsync(h->gstage, grad_host, (size_t)h->n * 8, cudaMemcpyHostToDevice, h->stream));
    int rc = ellcuda_update_begin(h, which, h->gstage, beta0, has_beta1, beta1);
    if (rc) return rc;
    if ((rc = ellcuda_update_finish(h))) return rc;
    return ellcuda_last_status(h, out_status);
}

double* ellcuda_grad_t_ptr(ell_handle h) { return h && h->kind == ELL_KIND_ELL ? h->xblk : nullptr; }   // null for a group
int32_t* ellcuda_status_ptr(ell_handle h) {
    if (h && !h->parts.empty()) h = h->parts[0];
    return !h ? nullptr : h->kind == ELL_KIND_ELLSTABLE ? h->st_status : &h->sc->status;
}

int ellcuda_last_status(ell_handle h, int32_t* out_status) {
    if (!h || !out_status) { set_error("null argument"); return ELLCUDA_E_ARG; }
    if (!h->parts.empty()) {   // every device is drained; the slab handles hold identical scalars, any mismatch is a lost peer
        int32_t st0 = ELL_UNKNOWN;
        for (size_t r = 0; r < h->parts.size(); ++r) {
            int32_t st = ELL_UNKNOWN;
            const int rc = ellcuda_last_status(h->parts[r], &st);
            if (rc) return rc;
            if (r == 0) st0 = st;
            else if (st != st0) st0 = ELL_UNKNOWN;
        }
        *out_status = st0;
        return ELLCUDA_OK;
    }
    ELL_DEVICE_SCOPE(h->device);
    const int* src = h->kind == ELL_KIND_ELLSTABLE ? h->st_status : &h->sc->status;
    ELL_CUDA_TRY(cudaMemcpyAsync(h->status_host, src, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    ELL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    *out_status = *h->status_host;
    return ELLCUDA_OK;
}

int ellcuda_exchange_export(ell_handle h, void* handle_out) {
    if (!h || !handle_out) { set_error("null argument"); return ELLCUDA_E_ARG; }
    if (!h->parts.empty()) { set_error("a group handle manages its own exchange"); return ELLCUDA_E_ARG; }
    if (h->kind != ELL_KIND_ELL) { set_error("EllStable does not shard"); return ELLCUDA_E_ARG; }
    ELL_DEVICE_SCOPE(h->device);
    cudaIpcMemHandle_t ipc;
    static_assert(sizeof(ipc) == ELLCUDA_IPC_HANDLE_BYTES, "IPC handle size");
    ELL_CUDA_TRY(cudaIpcGetMemHandle(&ipc, h->xblk));
    memcpy(handle_out, &ipc, sizeof ipc);
    return ELLCUDA_OK;
}

void* ellcuda_exchange_block_ptr(ell_handle h) { return h ? h->xblk : nullptr; }

static int attach_common(ell_handle h, int32_t nranks, int32_t rank) {
    if (!h->parts.empty()) { set_error("a group handle manages its own exchange"); return ELLCUDA_E_ARG; }
    if (h->sym && (h->sym_count != nranks || h->sym_rank != rank)) {
        set_error("a symmetric-storage handle attaches to exactly shard_count peers as rank shard_rank");
        return ELLCUDA_E_ARG;
    }
    if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) { set_error("bad nranks/rank"); return ELLCUDA_E_ARG; }
    if (h->nranks > 0) { set_error("exchange already attached"); return ELLCUDA_E_ARG; }
    ELL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return ELLCUDA_OK;
}

int ellcuda_exchange_attach_ptrs(ell_handle h, int32_t nranks, int32_t rank, void* const* blocks) {
    if (!h || !blocks) { set_error("null argument"); return ELLCUDA_E_ARG; }
    if (h->kind != ELL_KIND_ELL) { set_error("EllStable does not shard"); return ELLCUDA_E_ARG; }
    ELL_DEVICE_SCOPE(h->device);
    int rc = attach_common(h, nranks, rank);
    if (rc) return rc;
    for (int r = 0; r < nranks; ++r) {
        h->peer_blk[r] = r == rank ? h->xblk : (double*)blocks[r];
        if (!h->peer_blk[r]) { set_error("null peer block"); return ELLCUDA_E_ARG; }
        if (r == rank) continue;
        // a block on another device of this process: the sweep stores into it over NVLink, which needs peer access enabled
        cudaPointerAttributes pa;
        ELL_CUDA_TRY(cudaPointerGetAttributes(&pa, h->peer_blk[r]));
        if (pa.type == cudaMemoryTypeDevice && pa.device != h->device) {
            int can = 0;
            ELL_CUDA_TRY(cudaDeviceCanAccessPeer(&can, h->device, pa.device));
            if (!can) { set_error("no peer access between the devices of this exchange"); return ELLCUDA_E_COMM; }
            const cudaError_t e = cudaDeviceEnablePeerAccess(pa.device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__); return ELLCUDA_E_COMM; }
        }
    }
    h->nranks = nranks; h->rank = rank;
    return ELLCUDA_OK;
}

// attaches the slab handles of a group to each other (all in this process)
static int attach_group(ellcuda_ctx* g) {
    const int P = (int)g->parts.size();
    if (P < 2) return ELLCUDA_OK;
    void* blocks[kMaxRanks];
    for (int r = 0; r < P; ++r) blocks[r] = g->parts[r]->xblk;
    for (int r = 0; r < P; ++r) {
        const int rc = ellcuda_exchange_attach_ptrs(g->parts[r], P, r, blocks);
        if (rc) return rc;
    }
    return ELLCUDA_OK;
}

// Row-sharded Ell over `ndev` devices of THIS process (`ndev, devs`): one slab handle per device, rows split
// as evenly as possible, the exchange blocks attached by pointer with peer access.  The returned handle is used like any
// other Ell handle (ellcuda_update with a host gradient, xc, tsq, kappa, get_mq, dot_mv, clone, destroy); no launcher, no
// torch, no NCCL — a Rust host needs nothing but this library.
int ellcuda_create_sharded(const ell_create_desc* d, int32_t ndev, const int32_t* devs, ell_handle* out) {
    if (!d || !out || !devs) { set_error("null argument"); return ELLCUDA_E_ARG; }
    *out = nullptr;
    if (d->kind != ELL_KIND_ELL) { set_error("EllStable does not shard (its recurrences are triangular)"); return ELLCUDA_E_ARG; }
    if (ndev < 1 || ndev > kMaxRanks) { set_error("ndev must be in [1, 16]"); return ELLCUDA_E_ARG; }
    if (d->n < ndev) { set_error("fewer rows than devices"); return ELLCUDA_E_ARG; }
    if (d->row_begin != 0 || (d->row_end != 0 && d->row_end != d->n)) { set_error("ellcuda_create_sharded splits the rows itself"); return ELLCUDA_E_ARG; }
    // devs may repeat: several slabs then share a device, each on its own stream (how the tests run a group on one GPU)
    ellcuda_ctx* g = new ellcuda_ctx();
    g->kind = ELL_KIND_ELL; g->device = devs[0]; g->n = d->n; g->w = d->n;
    g->reduce_mode = d->reduce_mode; g->no_defer = d->no_defer_trick; g->use_parallel = d->use_parallel_cut;
    const long long base = d->n / ndev, rem = d->n % ndev;
    for (int r = 0; r < ndev; ++r) {
        ell_create_desc dr = *d;
        dr.device = devs[r];
        if (d->storage == ELL_STORAGE_SYMMETRIC) {   // column-block pairs dealt round-robin, the exchange is a reduction
            dr.row_begin = 0; dr.row_end = d->n;
            dr.shard_rank = r; dr.shard_count = ndev;
        } else {
            dr.row_begin = r * base + (r < rem ? r : rem);
            dr.row_end = dr.row_begin + base + (r < rem ? 1 : 0);
        }
        dr.stream = nullptr;   // every slab runs on its own stream of its own device ...
        for (int q = 0; q < r; ++q)   // ... but slabs that share a device share ONE stream: a finish kernel waits for the other
            if (devs[q] == devs[r]) { dr.stream = g->parts[q]->stream; break; }   // slabs' sweeps, and kernels that wait on one
        ell_handle c = nullptr;       // another must never be separate launches that could run in either order on one GPU
        const int rc = ellcuda_create(&dr, &c);
        if (rc) { free_ctx(g); return rc; }
        g->parts.push_back(c);
    }
    const int rc = attach_group(g);
    if (rc) { free_ctx(g); return rc; }
    *out = g;
    return ELLCUDA_OK;
}

int ellcuda_exchange_attach(ell_handle h, int32_t nranks, int32_t rank, const void* handles) {
    if (!h || !handles) { set_error("null argument"); return ELLCUDA_E_ARG; }
    if (h->kind != ELL_KIND_ELL) { set_error("EllStable does not shard"); return ELLCUDA_E_ARG; }
    ELL_DEVICE_SCOPE(h->device);
    int rc = attach_common(h, nranks, rank);
    if (rc) return rc;
    const char* hb = (const char*)handles;
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) { h->peer_blk[r] = h->xblk; continue; }
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, hb + (size_t)r * ELLCUDA_IPC_HANDLE_BYTES, sizeof ipc);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int q = 0; q < r; ++q)
                if (h->peer_opened[q]) { cudaIpcCloseMemHandle(h->peer_blk[q]); h->peer_opened[q] = false; h->peer_blk[q] = nullptr; }
            cuda_fail(e, "cudaIpcOpenMemHandle", __FILE__, __LINE__);
            return ELLCUDA_E_COMM;
        }
        h->peer_blk[r] = (double*)p;
        h->peer_opened[r] = true;
    }
    h->nranks = nranks; h->rank = rank;
    h->sym_fused = h->sym && nranks > 1;   // IPC peers are other processes on other GPUs: phase 2 of the reduction can run in the same kernel
    return ELLCUDA_OK;
}

int ellcuda_set_stream(ell_handle h, void* stream) {
    if (!h) { set_error("null handle"); return ELLCUDA_E_ARG; }
    if (!h->parts.empty()) { set_error("a group handle runs on its own per-device streams"); return ELLCUDA_E_ARG; }
    ELL_DEVICE_SCOPE(h->device);
    ELL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
    if (stream) h->stream = (cudaStream_t)stream;
    else {
        ELL_CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
    }
    return ELLCUDA_OK;
}

int ellcuda_sync(ell_handle h) {
    if (!h) { set_error("null handle"); return ELLCUDA_E_ARG; }
    if (!h->parts.empty()) {
        for (ellcuda_ctx* p : h->parts) { const int rc = ellcuda_sync(p); if (rc) return rc; }
        return ELLCUDA_OK;
    }
    ELL_DEVICE_SCOPE(h->device);
    ELL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return ELLCUDA_OK;
}

int ellcuda_set_timing(ell_handle h, int32_t on) {
    if (!h) { set_error("null handle"); return ELLCUDA_E_ARG; }
    h->timing = on != 0;
    h->timed = false;
    for (ellcuda_ctx* p : h->parts) { p->timing = h->timing; p->timed = false; }
    return ELLCUDA_OK;
}

float ellcuda_last_sweep_ms(ell_handle h) {
    if (h && !h->parts.empty()) {   // group: the slowest slab
        float worst = -1.0f;
        for (ellcuda_ctx* p : h->parts) { const float ms = ellcuda_last_sweep_ms(p); if (ms > worst) worst = ms; }
        return worst;
    }
    if (!h || !h->timed) return -1.0f;
    float ms = -1.0f;
    if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.0f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.0f;
    return ms;
}

int64_t ellcuda_handle_launch_count(ell_handle h) {
    if (!h) return 0;
    long long total = h->launches;
    for (ellcuda_ctx* p : h->parts) total += p->launches;
    return total;
}
int32_t ellcuda_group_size(ell_handle h) { return !h ? 0 : h->parts.empty() ? 1 : (int32_t)h->parts.size(); }

}  // extern "C"
#endif  // __CUDACC__
