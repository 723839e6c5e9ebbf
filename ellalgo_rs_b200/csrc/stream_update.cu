// stream_update.cu — large-n mode: ONE ellipsoid whose n×n matrix lives in HBM (src/ell.rs:9-16, 97-137).
//
// The reference streams M three times per update (GEMV read, rank-1 read + write).  Here one update costs ONE read
// and ONE write of M (16·n² bytes, the algorithmic minimum), by deferring the rank-1 of update k into the sweep
// that forms grad_t of update k+1:
//
//     sweep(k+1):   M <- (M - (ratio_k·g̃_k[hi])·g̃_k[lo]) [·kappa if no_defer_trick]     src/ell.rs:117-128, 132-135
//                   g̃_{k+1}[c] = sum_i M[i][c]·g_{k+1}[i]   (i ascending)                src/arr.rs:426-442
//     finish(k+1):  omega, tsq, EllCalc, xc, kappa                                         src/ell.rs:103-115, 130
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
#include <vector>

#include "common.cuh"

namespace ellcuda {
namespace {

constexpr int kCols = 64;    // columns per warp (2 per lane)
constexpr int kRowsDefault = 64;    // rows per TMA tile   (measured at n=16384: 64x3 682 us, 32x6 693, 48x4 692,
constexpr int kStagesDefault = 3;   // smem ring depth      32x4 709, 48x3 753, 32x3 840, 16x8 831, 64x4 1219 [1 CTA/SM])
constexpr int kRowsMax = 64;
// per tile: kCols*ROWS*8 bytes of M and ROWS*32 bytes of pack {g[i], gt[i], ratio*gt[i], 0}

struct Scal {  // device-resident scalars of the Ell (src/ell.rs:9-16) + the deferred rank-1
    double kappa;
    double tsq;
    double omega;
    double ratio;   // sigma/omega of the pending rank-1
    double scale;   // kappa to fold into M with the pending rank-1 (no_defer_trick), else 1
    int pending;    // 1: (ratio, scale, gt) describe a rank-1 not yet applied to M
    int status;     // CutStatus of the last finish
    unsigned long long epoch;   // peer exchange: number of the update in flight (device-resident so that a captured
                                // CUDA graph of one update can be replayed); bumped at the end of finish
};

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, void* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, const void* src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1),
                 "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ kernels
// pack[i] = {g[i], gt[i], ratio·gt[i], 0} for i < n, zeros up to npad; also keeps a copy of g for `finish`.
__global__ void pack_kernel(const double* __restrict__ g_in, double* __restrict__ g_keep, const double* __restrict__ gt_block,
                            int attached, const Scal* __restrict__ sc, double4* __restrict__ pack, long long n, long long npad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    // grad_t of the previous update: with peers attached it sits in the buffer of the other parity
    const double* gt = attached ? gt_block + ((sc->epoch & 1ULL) ^ 1ULL) * n : gt_block;
    double4 v = make_double4(0.0, 0.0, 0.0, 0.0);
    if (i < n) {
        const double g = g_in ? g_in[i] : 0.0;
        g_keep[i] = g;
        v.x = g;
        if (sc->pending) {
            const double t = gt[i];
            v.y = t;
            v.z = sc->ratio * t;  // r_qg = ratio * grad_t[i]   (src/ell.rs:119)
        }
    }
    pack[i] = v;
}

// MODE 0: every row of the tile is above the block's columns (i < c):   update = (ratio·gt[c]) · gt[i]
// MODE 1: every row is at or below them (i >= c):                        update = (ratio·gt[i]) · gt[c]
// MODE 2: the tile crosses the diagonal: choose per element.             (src/ell.rs:118-127: hi = max, lo = min)
template <int MODE, bool PENDING, bool SCALE>
__device__ __forceinline__ void tile_rows(double2* __restrict__ tile, const double4* __restrict__ pk, int lane, int nrows, int row0,
                                          int col0, double gt_c0, double gt_c1, double rg_c0, double rg_c1, double scale, double& acc0,
                                          double& acc1) {
#pragma unroll 8
    for (int r = 0; r < nrows; ++r) {  // nrows < kRows only in the last tile of a segment (rows past n are not M)
        double2 m = tile[r * (kCols / 2) + lane];
        const double4 p = pk[r];  // broadcast
        if (PENDING) {
            double u0, u1;
            if (MODE == 0) {
                u0 = rg_c0 * p.y;
                u1 = rg_c1 * p.y;
            } else if (MODE == 1) {
                u0 = p.z * gt_c0;
                u1 = p.z * gt_c1;
            } else {
                const int i = row0 + r;
                u0 = (i >= col0) ? p.z * gt_c0 : rg_c0 * p.y;
                u1 = (i >= col0 + 1) ? p.z * gt_c1 : rg_c1 * p.y;
            }
            m.x -= u0;
            m.y -= u1;
            if (SCALE) {
                m.x *= scale;  // mq *= kappa   (src/ell.rs:132-135)
                m.y *= scale;
            }
            tile[r * (kCols / 2) + lane] = m;
        }
        acc0 += m.x * p.x;  // grad_t[c] += M[i][c] * g[i], i ascending
        acc1 += m.y * p.x;
    }
}

template <int kRows, int kStages, bool PENDING, bool SCALE>
__device__ __forceinline__ void sweep_body(const CUtensorMap* map, const double4* __restrict__ pack, double* __restrict__ out,
                                           long long w, long long row_lo, long long row_hi, long long c_glob0, double scale,
                                           unsigned char* smem_raw) {
    constexpr int kTileBytes = kCols * kRows * 8;
    constexpr int kPackBytes = kRows * 32;
    // smem: [kStages tiles][kStages packs][kStages mbarriers]
    double2* tiles = reinterpret_cast<double2*>(smem_raw);
    double4* packs = reinterpret_cast<double4*>(smem_raw + kStages * kTileBytes);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + kStages * (kTileBytes + kPackBytes));
    const int lane = threadIdx.x;
    const int cb = blockIdx.x;                 // column block of this warp
    const long long c_loc0 = (long long)cb * kCols;  // first local column
    const int ntiles = (int)((row_hi - row_lo + kRows - 1) / kRows);

    if (lane == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    auto issue = [&](int t) {  // lane 0 only
        const int s = t % kStages;
        const long long r0 = row_lo + (long long)t * kRows;
        mbar_expect_tx(&bars[s], kTileBytes + kPackBytes);
        tma_load_2d(reinterpret_cast<unsigned char*>(tiles) + s * kTileBytes, map, (int)c_loc0, (int)r0, &bars[s]);
        bulk_load_1d(reinterpret_cast<unsigned char*>(packs) + s * kPackBytes, pack + r0, kPackBytes, &bars[s]);
    };
    if (lane == 0) {
        const int pre = ntiles < (kStages - 1) ? ntiles : (kStages - 1);
        for (int t = 0; t < pre; ++t) issue(t);
    }

    // this lane's two columns (global indices) and their pending-vector entries
    const long long cg0 = c_glob0 + c_loc0 + 2 * lane;
    double gt_c0 = 0.0, gt_c1 = 0.0, rg_c0 = 0.0, rg_c1 = 0.0;
    if (PENDING) {
        // pack is indexed by global row = global column index of the symmetric matrix
        const double4 a = pack[cg0];
        const double4 b = pack[cg0 + 1];  // pack is padded, always readable
        gt_c0 = a.y; rg_c0 = a.z;
        gt_c1 = b.y; rg_c1 = b.z;
    }
    double acc0 = 0.0, acc1 = 0.0;
    const long long blk_lo = c_glob0 + c_loc0, blk_hi = blk_lo + kCols;  // global column range of the block

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % kStages;
        mbar_wait(&bars[s], (unsigned)((t / kStages) & 1));
        double2* tile = reinterpret_cast<double2*>(reinterpret_cast<unsigned char*>(tiles) + s * kTileBytes);
        const double4* pk = reinterpret_cast<const double4*>(reinterpret_cast<unsigned char*>(packs) + s * kPackBytes);
        const long long r0 = row_lo + (long long)t * kRows;
        const int nrows = (int)((row_hi - r0) < kRows ? (row_hi - r0) : kRows);
        if (r0 + kRows <= blk_lo)
            tile_rows<0, PENDING, SCALE>(tile, pk, lane, nrows, 0, 0, gt_c0, gt_c1, rg_c0, rg_c1, scale, acc0, acc1);
        else if (r0 >= blk_hi - 1)
            tile_rows<1, PENDING, SCALE>(tile, pk, lane, nrows, 0, 0, gt_c0, gt_c1, rg_c0, rg_c1, scale, acc0, acc1);
        else
            tile_rows<2, PENDING, SCALE>(tile, pk, lane, nrows, (int)(r0 - blk_lo), 2 * lane, gt_c0, gt_c1, rg_c0, rg_c1, scale,
                                         acc0, acc1);
        if (PENDING) fence_async_smem();  // generic-proxy writes -> visible to the TMA store
        __syncwarp();
        if (lane == 0) {
            if (PENDING) {
                tma_store_2d(map, (int)c_loc0, (int)r0, tile);
                bulk_commit();
            }
            const int nt = t + kStages - 1;  // refill the stage whose store was issued one iteration ago
            if (nt < ntiles) {
                if (PENDING) bulk_wait_read<1>();
                issue(nt);
            }
        }
        __syncwarp();
    }
    if (PENDING && lane == 0) bulk_wait_all();
    const long long cl = c_loc0 + 2 * lane;
    if (cl < w) out[cl] = acc0;
    if (cl + 1 < w) out[cl + 1] = acc1;
}

// grid = (column blocks, row segments); block = 32 threads.
template <int kRows, int kStages>
__global__ void __launch_bounds__(32) sweep_kernel(const __grid_constant__ CUtensorMap map, const double4* __restrict__ pack,
                                                   const Scal* __restrict__ sc, double* __restrict__ partial, long long ld_part,
                                                   long long n, long long w, long long c_glob0, int nseg) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int seg = blockIdx.y;
    const long long rows_per = ((n + nseg - 1) / nseg + kRows - 1) / kRows * kRows;
    const long long row_lo = (long long)seg * rows_per;
    long long row_hi = row_lo + rows_per;
    if (row_hi > n) row_hi = n;
    double* out = partial + (long long)seg * ld_part;
    if (row_lo >= row_hi) {  // empty segment: contributes exact zeros
        const long long cl = (long long)blockIdx.x * kCols + 2 * threadIdx.x;
        if (cl < w) out[cl] = 0.0;
        if (cl + 1 < w) out[cl + 1] = 0.0;
        return;
    }
    const int pending = sc->pending;
    const double scale = sc->scale;
    if (!pending) sweep_body<kRows, kStages, false, false>(&map, pack, out, w, row_lo, row_hi, c_glob0, 1.0, smem_raw);
    else if (scale == 1.0) sweep_body<kRows, kStages, true, false>(&map, pack, out, w, row_lo, row_hi, c_glob0, 1.0, smem_raw);
    else sweep_body<kRows, kStages, true, true>(&map, pack, out, w, row_lo, row_hi, c_glob0, scale, smem_raw);
}

// grad_t[c_glob0 + c] = sum over row segments, in segment order (one segment: a copy).
__global__ void combine_kernel(const double* __restrict__ partial, long long ld_part, int nseg, double* __restrict__ gt,
                               long long w, long long c_glob0) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= w) return;
    double acc = partial[c];
    for (int s = 1; s < nseg; ++s) acc += partial[(long long)s * ld_part + c];
    gt[c_glob0 + c] = acc;
}

// Peer exchange (row-sharded handles attached with ellcuda_exchange_attach): the kernel that completes this rank's
// slab of grad_t stores it straight into EVERY rank's exchange block over NVLink (peer pointers), then publishes the
// update's epoch in each peer's flag word; `finish_kernel` on every rank waits for all P flags.  No NCCL call, no host
// involvement.  The block holds two grad_t buffers (epoch parity) so that a rank one update ahead never overwrites
// data a slower rank still reads (see the ordering argument in DESIGN.md §5).
constexpr int kMaxRanks = 16;
struct PeerTable {
    double* blk[kMaxRanks];   // rank r's exchange block: grad_t[2][n], then kMaxRanks flag words (index = source rank)
    int nranks;
    int rank;
};

__global__ void __launch_bounds__(1024) combine_push_kernel(const double* __restrict__ partial, long long ld_part, int nseg,
                                                            long long n, long long w, long long c_glob0, PeerTable pt,
                                                            const Scal* __restrict__ sc) {
    const unsigned long long epoch = sc->epoch;
    const long long off = (long long)(epoch & 1ULL) * n + c_glob0;
    for (long long c = threadIdx.x; c < w; c += blockDim.x) {
        double acc = partial[c];
        for (int s = 1; s < nseg; ++s) acc += partial[(long long)s * ld_part + c];
        for (int r = 0; r < pt.nranks; ++r) pt.blk[r][off + c] = acc;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < pt.nranks) {
        __threadfence_system();
        unsigned long long* flag = reinterpret_cast<unsigned long long*>(pt.blk[threadIdx.x] + 2 * n) + pt.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(epoch) : "memory");
    }
}

// returns false on timeout (~2 s) so that a lost peer cannot hang the GPU
__device__ __forceinline__ bool wait_flag(const unsigned long long* flag, unsigned long long epoch) {
    const long long t0 = clock64();
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        if (v >= epoch) return true;
        if (clock64() - t0 > 4000000000LL) return false;
        __nanosleep(200);
    }
}

// omega, tsq, EllCalc, xc, kappa — one CTA, identical on every rank (src/ell.rs:103-115, 130-135).
constexpr int kFinThreads = 1024;
constexpr int kFinChunk = 4096;  // products staged per pass (32 KB of shared memory)
__global__ void __launch_bounds__(kFinThreads) finish_kernel(const double* __restrict__ g, const double* gt_block,
                                                             double* __restrict__ xc, Scal* __restrict__ sc, long long n,
                                                             CalcConsts cc, int which, Cut cut, int no_defer, int reduce_mode,
                                                             int wait_n) {
    __shared__ double prod[kFinChunk];
    __shared__ double s_omega;
    __shared__ double s_roo;
    __shared__ int s_status;
    __shared__ int s_timeout;
    const int tid = threadIdx.x;
    const unsigned long long epoch = sc->epoch;
    const double* gt = wait_n > 0 ? gt_block + (epoch & 1ULL) * n : gt_block;
    if (wait_n > 0) {   // peer exchange: all P slabs of grad_t must have landed in this rank's block
        const unsigned long long* wait_flags = reinterpret_cast<const unsigned long long*>(gt_block + 2 * n);
        if (tid == 0) s_timeout = 0;
        __syncthreads();
        if (tid < wait_n && !wait_flag(wait_flags + tid, epoch)) s_timeout = 1;
        __syncthreads();
        if (s_timeout) {
            if (tid == 0) { sc->status = kUnknown; sc->pending = 0; sc->epoch = epoch + 1; }
            return;
        }
    }
    if (reduce_mode == 0) {
        // grad.dot(&grad_t): products in parallel, the sum strictly left to right (src/arr.rs:443-451)
        double acc = 0.0;
        for (long long base = 0; base < n; base += kFinChunk) {
            const int len = (int)((n - base) < kFinChunk ? (n - base) : kFinChunk);
            for (int k = tid; k < len; k += kFinThreads) prod[k] = g[base + k] * gt[base + k];
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
        // blocked: 1024 ordered chunks summed left to right, then a fixed binary tree — deterministic
        const long long per = (n + kFinThreads - 1) / kFinThreads;
        const long long lo = (long long)tid * per;
        long long hi = lo + per;
        if (hi > n) hi = n;
        double acc = 0.0;
        for (long long k = lo; k < hi; ++k) acc += g[k] * gt[k];
        prod[tid] = acc;
        __syncthreads();
        for (int off = kFinThreads / 2; off > 0; off >>= 1) {
            if (tid < off) prod[tid] = prod[tid] + prod[tid + off];
            __syncthreads();
        }
        if (tid == 0) s_omega = prod[0];
    }
    __syncthreads();
    if (tid == 0) {
        const double omega = s_omega;
        const double tsq = sc->kappa * omega;
        sc->tsq = tsq;
        sc->omega = omega;
        Rsd r;
        const int status = ell_calc(cc, which, cut, tsq, r);
        sc->status = status;
        s_status = status;
        if (status == kSuccess) {
            s_roo = r.rho / omega;
            sc->ratio = r.sigma / omega;
            double kappa = sc->kappa * r.delta;
            if (no_defer) {
                sc->scale = kappa;
                kappa = 1.0;
            } else {
                sc->scale = 1.0;
            }
            sc->kappa = kappa;
            sc->pending = 1;
        } else {
            sc->pending = 0;  // state unchanged except tsq (src/ell.rs:107-109)
        }
    }
    __syncthreads();
    if (s_status == kSuccess) {
        const double roo = s_roo;
        for (long long i = tid; i < n; i += kFinThreads) xc[i] -= roo * gt[i];
    }
    if (tid == 0 && wait_n > 0) sc->epoch = epoch + 1;   // every thread has read `epoch` before the first barrier
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

// ------------------------------------------------------------------------------------------------ tensor map
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tensor_map(CUtensorMap* map, void* base, long long rows, long long cols, long long ld, int box_rows) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        ELL_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled not available"); return ELLCUDA_E_CUDA; }
        fn = (EncodeTiledFn)p;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    const cuuint32_t box[2] = {(cuuint32_t)kCols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    if (const char* e = getenv("ELLCUDA_L2PROMO")) {   // performance experiments
        const int v = atoi(e);
        promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
              : v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    }
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return ELLCUDA_E_CUDA;
    }
    return ELLCUDA_OK;
}

inline size_t xblk_bytes(long long n) { return (size_t)2 * n * 8 + kMaxRanks * 8; }
inline unsigned long long* xblk_flags(double* blk, long long n) { return reinterpret_cast<unsigned long long*>(blk + 2 * n); }
constexpr size_t sweep_smem(int rows, int stages) { return (size_t)stages * (kCols * rows * 8 + rows * 32) + stages * 8 + 128; }

}  // namespace
}  // namespace ellcuda

// =================================================================================================== handle
struct ellcuda_ctx {
    int device = 0;
    long long n = 0, w = 0, c0 = 0, ld = 0, npad = 0, ld_part = 0;
    int nseg = 1, reduce_mode = 0, no_defer = 0, use_parallel = 1;
    int tile_rows = ellcuda::kRowsDefault, stages = ellcuda::kStagesDefault;   // sweep tile shape (ELLCUDA_SWEEP=rows,stages)
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    double* S = nullptr;        // n × ld, column block [c0, c0+w) of the symmetric matrix
    double* xc = nullptr;       // n
    double* g = nullptr;        // n, gradient of the update in flight
    double* gstage = nullptr;   // n, landing buffer for host gradients (ellcuda_update)
    int* status_host = nullptr; // pinned, CutStatus read-back
    double* xblk = nullptr;     // exchange block: grad_t[2][n] then kMaxRanks flag words (IPC-shareable cudaMalloc)
    double* gt = nullptr;       // n, current grad_t buffer inside xblk (full length; a rank writes [c0, c0+w))
    int nranks = 0, rank = 0;                // > 0 once ellcuda_exchange_attach* succeeded
    double* peer_blk[ellcuda::kMaxRanks] = {};   // every rank's exchange block as mapped in this process
    bool peer_opened[ellcuda::kMaxRanks] = {};   // mapped by cudaIpcOpenMemHandle (must be closed)
    double4* pack = nullptr;    // npad + kCols + 2 (tail stays zero)
    double* partial = nullptr;  // nseg × ld_part
    ellcuda::Scal* sc = nullptr;
    CUtensorMap map;
    int which = 0;
    ellcuda::Cut cut{0.0, 0.0, false};
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    long long launches = 0;
};

namespace ellcuda {
namespace {

void free_ctx(ellcuda_ctx* h) {
    if (!h) return;
    DevGuard guard(h->device);
    cudaFree(h->S); cudaFree(h->xc); cudaFree(h->g); cudaFree(h->gstage); cudaFreeHost(h->status_host); cudaFree(h->pack); cudaFree(h->partial); cudaFree(h->sc);
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
    h->ld = (h->w + 1) & ~1LL;  // TMA needs a 16-byte row pitch
    if (const char* e = getenv("ELLCUDA_SWEEP")) {   // performance experiments: "rows,stages"
        int r = 0, st = 0;
        if (sscanf(e, "%d,%d", &r, &st) == 2 && (r == 16 || r == 32 || r == 48 || r == 64) && (st == 3 || st == 4 || st == 6 || st == 8)) {
            h->tile_rows = r;
            h->stages = st;
        }
    }
    h->npad = (n + kRowsMax - 1) / kRowsMax * kRowsMax + kRowsMax;
    h->ld_part = (h->w + kCols - 1) / kCols * kCols;
    ELL_CUDA_TRY(cudaMalloc(&h->S, (size_t)n * h->ld * 8));
    ELL_CUDA_TRY(cudaMalloc(&h->xc, (size_t)n * 8));
    ELL_CUDA_TRY(cudaMalloc(&h->g, (size_t)n * 8));
    ELL_CUDA_TRY(cudaMalloc(&h->gstage, (size_t)n * 8));
    ELL_CUDA_TRY(cudaMallocHost(&h->status_host, sizeof(int)));
    keep_mempool(h->device);
    ELL_CUDA_TRY(cudaMalloc(&h->xblk, xblk_bytes(n)));
    ELL_CUDA_TRY(cudaMemset(h->xblk, 0, xblk_bytes(n)));
    h->gt = h->xblk;
    ELL_CUDA_TRY(cudaMalloc(&h->pack, (size_t)(h->npad + kCols + 2) * 32));
    ELL_CUDA_TRY(cudaMalloc(&h->partial, (size_t)h->nseg * h->ld_part * 8));
    ELL_CUDA_TRY(cudaMalloc(&h->sc, sizeof(Scal)));
    ELL_CUDA_TRY(cudaEventCreate(&h->ev0));
    ELL_CUDA_TRY(cudaEventCreate(&h->ev1));
    ELL_CUDA_TRY(cudaMemset(h->pack, 0, (size_t)(h->npad + kCols + 2) * 32));
    ELL_CUDA_TRY(cudaMemset(h->g, 0, (size_t)n * 8));
    int rc = make_tensor_map(&h->map, h->S, n, h->w, h->ld, h->tile_rows);
    if (rc) return rc;
    return ELLCUDA_OK;
}

template <int R, int ST>
int launch_sweep_t(ellcuda_ctx* h, dim3 grid, cudaStream_t s) {
    static bool attr_set = false;   // per instantiation
    constexpr size_t smem = sweep_smem(R, ST);
    if (!attr_set) {
        ELL_CUDA_TRY(cudaFuncSetAttribute(sweep_kernel<R, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    sweep_kernel<R, ST><<<grid, 32, smem, s>>>(h->map, h->pack, h->sc, h->partial, h->ld_part, h->n, h->w, h->c0, h->nseg);
    return ELLCUDA_OK;
}

int launch_sweep(ellcuda_ctx* h, dim3 grid, cudaStream_t s) {   // tile rows x ring depth (ELLCUDA_SWEEP, experiments)
    const int key = h->tile_rows * 100 + h->stages;
    switch (key) {
        case 1608: return launch_sweep_t<16, 8>(h, grid, s);
        case 3204: return launch_sweep_t<32, 4>(h, grid, s);
        case 3206: return launch_sweep_t<32, 6>(h, grid, s);
        case 4804: return launch_sweep_t<48, 4>(h, grid, s);
        default: return launch_sweep_t<64, 3>(h, grid, s);
    }
}

// begin = pack + sweep + combine (+ push to peers).  g_dev may be null (flush: applies the pending rank-1 only).
int enqueue_begin(ellcuda_ctx* h, const double* g_dev) {
    cudaStream_t s = h->stream;
    int rc_launch = 0;
    const int thr = 256;
    const bool flush = g_dev == nullptr;
    const bool peers = h->nranks > 0 && !flush;
    // with peers attached grad_t alternates between the block's two buffers (parity of the device-side epoch)
    pack_kernel<<<(unsigned)((h->npad + thr - 1) / thr), thr, 0, s>>>(g_dev, h->g, h->xblk, h->nranks > 0 ? 1 : 0, h->sc, h->pack,
                                                                      h->n, h->npad);
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &cap);
    const bool rec = h->timing;
    // inside a stream capture the records become external event-record nodes, so a replayed graph still times its sweeps
    const unsigned evflag = cap == cudaStreamCaptureStatusNone ? cudaEventRecordDefault : cudaEventRecordExternal;
    if (rec) ELL_CUDA_TRY(cudaEventRecordWithFlags(h->ev0, s, evflag));
    dim3 grid((unsigned)(h->ld_part / kCols), (unsigned)h->nseg);
    if ((rc_launch = launch_sweep(h, grid, s))) return rc_launch;
    if (rec) { ELL_CUDA_TRY(cudaEventRecordWithFlags(h->ev1, s, evflag)); h->timed = true; }
    if (peers) {
        PeerTable pt;
        pt.nranks = h->nranks;
        pt.rank = h->rank;
        for (int r = 0; r < h->nranks; ++r) pt.blk[r] = h->peer_blk[r];
        combine_push_kernel<<<1, 1024, 0, s>>>(h->partial, h->ld_part, h->nseg, h->n, h->w, h->c0, pt, h->sc);
    } else if (!flush) {
        combine_kernel<<<(unsigned)((h->w + thr - 1) / thr), thr, 0, s>>>(h->partial, h->ld_part, h->nseg, h->gt, h->w, h->c0);
    }
    ELL_CUDA_TRY(cudaGetLastError());
    h->launches += flush ? 2 : 3;
    return ELLCUDA_OK;
}

__global__ void clear_pending_kernel(Scal* sc) { sc->pending = 0; sc->scale = 1.0; }

int flush_pending(ellcuda_ctx* h) {  // make S hold the reference's `mq` exactly
    int rc = enqueue_begin(h, nullptr);
    if (rc) return rc;
    clear_pending_kernel<<<1, 1, 0, h->stream>>>(h->sc);
    ELL_CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return ELLCUDA_OK;
}

}  // namespace
}  // namespace ellcuda

using namespace ellcuda;

extern "C" {

int ellcuda_create(const ell_create_desc* d, ell_handle* out) {
    if (!d || !out) { set_error("null argument"); return ELLCUDA_E_ARG; }
    *out = nullptr;
    if (d->kind != ELL_KIND_ELL) { set_error("large-n mode exists for Ell only (EllStable does not shard; use batched mode)"); return ELLCUDA_E_ARG; }
    if (d->n <= 0 || !d->xc) { set_error("n must be positive and xc given"); return ELLCUDA_E_ARG; }
    long long rb = d->row_begin, re = d->row_end;
    if (rb == 0 && re == 0) re = d->n;
    if (rb < 0 || re > d->n || rb >= re) { set_error("bad row range"); return ELLCUDA_E_ARG; }
    if (d->reduce_mode < 0 || d->reduce_mode > 64) { set_error("reduce_mode must be in [0, 64]"); return ELLCUDA_E_ARG; }
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
    if (d->mq) {  // symmetric input required (see capi.cu)
        const long long n = d->n;
        for (long long i = 0; i < n; ++i)
            for (long long j = 0; j < i; ++j)
                if (!(d->mq[i * n + j] == d->mq[j * n + i])) { set_error("mq must be symmetric"); return ELLCUDA_E_ARG; }
    }
    ellcuda_ctx* h = new ellcuda_ctx();
    h->device = d->device;
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
    if (d->mq) {  // column block [c0, c0+w) of the row-major host matrix
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

int ellcuda_clone(ell_handle h, ell_handle* out) {
    if (!h || !out) { set_error("null handle"); return ELLCUDA_E_ARG; }
    ELL_DEVICE_SCOPE(h->device);
    ellcuda_ctx* c = new ellcuda_ctx();
    c->device = h->device; c->n = h->n; c->w = h->w; c->c0 = h->c0;
    c->nseg = h->nseg; c->reduce_mode = h->reduce_mode; c->no_defer = h->no_defer; c->use_parallel = h->use_parallel;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
    c->own_stream = true;
    int rc = alloc_ctx(c);
    if (rc) { free_ctx(c); return rc; }
    cudaStream_t s = h->stream;
    const long long n = h->n;
    e = cudaMemcpyAsync(c->S, h->S, (size_t)n * h->ld * 8, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->xc, h->xc, (size_t)n * 8, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->g, h->g, (size_t)n * 8, cudaMemcpyDeviceToDevice, s);
    // the clone is detached from peers: it keeps grad_t of the last completed update in buffer 0 and restarts its epoch
    Scal sc_h;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&sc_h, h->sc, sizeof(Scal), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    const double* src_gt = h->nranks > 0 ? h->xblk + ((sc_h.epoch & 1ULL) ^ 1ULL) * n : h->xblk;
    sc_h.epoch = 1;
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->gt, src_gt, (size_t)n * 8, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->sc, &sc_h, sizeof(Scal), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { free_ctx(c); return cuda_fail(e, "ellcuda_clone copy", __FILE__, __LINE__); }
    *out = c;
    return ELLCUDA_OK;
}

void ellcuda_destroy(ell_handle h) { free_ctx(h); }

int ellcuda_xc(ell_handle h, double* out) {
    if (!h || !out) { set_error("null argument"); return ELLCUDA_E_ARG; }
    ELL_DEVICE_SCOPE(h->device);
    ELL_CUDA_TRY(cudaMemcpyAsync(out, h->xc, (size_t)h->n * 8, cudaMemcpyDeviceToHost, h->stream));
    ELL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return ELLCUDA_OK;
}

int ellcuda_set_xc(ell_handle h, const double* x) {
    if (!h || !x) { set_error("null argument"); return ELLCUDA_E_ARG; }
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

int ellcuda_tsq(ell_handle h, double* out) {
    if (!h || !out) { set_error("null argument"); return ELLCUDA_E_ARG; }
    Scal sc;
    int rc = read_scal(h, &sc);
    if (rc) return rc;
    *out = sc.tsq;
    return ELLCUDA_OK;
}

int ellcuda_kappa(ell_handle h, double* out) {
    if (!h || !out) { set_error("null argument"); return ELLCUDA_E_ARG; }
    Scal sc;
    int rc = read_scal(h, &sc);
    if (rc) return rc;
    *out = sc.kappa;
    return ELLCUDA_OK;
}

int ellcuda_get_mq(ell_handle h, double* out) {  // this slab's rows, w × n row-major
    if (!h || !out) { set_error("null argument"); return ELLCUDA_E_ARG; }
    ELL_DEVICE_SCOPE(h->device);
    Scal sc;
    int rc = read_scal(h, &sc);
    if (rc) return rc;
    if (sc.pending && (rc = flush_pending(h))) return rc;
    double* tmp = nullptr;
    ELL_CUDA_TRY(cudaMallocAsync((void**)&tmp, (size_t)h->w * h->n * 8, h->stream));
    dim3 grid((unsigned)((h->w + 31) / 32), (unsigned)((h->n + 31) / 32)), block(32, 8);
    transpose_kernel<<<grid, block, 0, h->stream>>>(h->S, h->ld, tmp, h->n, h->w);
    h->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, tmp, (size_t)h->w * h->n * 8, cudaMemcpyDeviceToHost, h->stream);
    cudaFreeAsync(tmp, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return cuda_fail(e, "ellcuda_get_mq", __FILE__, __LINE__);
    return ELLCUDA_OK;
}

int ellcuda_update_begin(ell_handle h, int32_t which, const double* grad_dev, double beta0, int32_t has_beta1, double beta1) {
    if (!h || !grad_dev) { set_error("null argument"); return ELLCUDA_E_ARG; }
    if (which < ELL_UPDATE_BIAS || which > ELL_UPDATE_Q) { set_error("bad update selector"); return ELLCUDA_E_ARG; }
    ELL_DEVICE_SCOPE(h->device);
    h->which = which;
    h->cut = has_beta1 > 0 ? parallel_cut(beta0, beta1) : single_cut(beta0);
    return enqueue_begin(h, grad_dev);
}

int ellcuda_update_finish(ell_handle h) {
    if (!h) { set_error("null handle"); return ELLCUDA_E_ARG; }
    ELL_DEVICE_SCOPE(h->device);
    const CalcConsts cc = CalcConsts::make((int)h->n, h->use_parallel);
    const bool peers = h->nranks > 0;
    finish_kernel<<<1, kFinThreads, 0, h->stream>>>(h->g, h->xblk, h->xc, h->sc, h->n, cc, h->which, h->cut, h->no_defer, h->reduce_mode,
                                                    peers ? h->nranks : 0);
    ELL_CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return ELLCUDA_OK;
}

int ellcuda_update(ell_handle h, int32_t which, const double* grad_host, double beta0, int32_t has_beta1, double beta1,
                   int32_t* out_status) {
    if (!h || !grad_host || !out_status) { set_error("null argument"); return ELLCUDA_E_ARG; }
    if (h->w != h->n) { set_error("ellcuda_update needs a whole-matrix handle; sharded handles use update_begin/finish"); return ELLCUDA_E_ARG; }
    ELL_DEVICE_SCOPE(h->device);
    ELL_CUDA_TRY(cudaMemcpyAsync(h->gstage, grad_host, (size_t)h->n * 8, cudaMemcpyHostToDevice, h->stream));
    int rc = ellcuda_update_begin(h, which, h->gstage, beta0, has_beta1, beta1);
    if (rc) return rc;
    if ((rc = ellcuda_update_finish(h))) return rc;
    return ellcuda_last_status(h, out_status);
}

double* ellcuda_grad_t_ptr(ell_handle h) { return h ? h->gt : nullptr; }
int32_t* ellcuda_status_ptr(ell_handle h) { return h ? &h->sc->status : nullptr; }

int ellcuda_last_status(ell_handle h, int32_t* out_status) {
    if (!h || !out_status) { set_error("null argument"); return ELLCUDA_E_ARG; }
    ELL_DEVICE_SCOPE(h->device);
    ELL_CUDA_TRY(cudaMemcpyAsync(h->status_host, &h->sc->status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    ELL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    *out_status = *h->status_host;
    return ELLCUDA_OK;
}

int ellcuda_exchange_export(ell_handle h, void* handle_out) {
    if (!h || !handle_out) { set_error("null argument"); return ELLCUDA_E_ARG; }
    ELL_DEVICE_SCOPE(h->device);
    cudaIpcMemHandle_t ipc;
    static_assert(sizeof(ipc) == ELLCUDA_IPC_HANDLE_BYTES, "IPC handle size");
    ELL_CUDA_TRY(cudaIpcGetMemHandle(&ipc, h->xblk));
    memcpy(handle_out, &ipc, sizeof ipc);
    return ELLCUDA_OK;
}

void* ellcuda_exchange_block_ptr(ell_handle h) { return h ? h->xblk : nullptr; }

static int attach_common(ell_handle h, int32_t nranks, int32_t rank) {
    if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) { set_error("bad nranks/rank"); return ELLCUDA_E_ARG; }
    if (h->nranks > 0) { set_error("exchange already attached"); return ELLCUDA_E_ARG; }
    ELL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return ELLCUDA_OK;
}

int ellcuda_exchange_attach_ptrs(ell_handle h, int32_t nranks, int32_t rank, void* const* blocks) {
    if (!h || !blocks) { set_error("null argument"); return ELLCUDA_E_ARG; }
    ELL_DEVICE_SCOPE(h->device);
    int rc = attach_common(h, nranks, rank);
    if (rc) return rc;
    for (int r = 0; r < nranks; ++r) {
        h->peer_blk[r] = r == rank ? h->xblk : (double*)blocks[r];
        if (!h->peer_blk[r]) { set_error("null peer block"); return ELLCUDA_E_ARG; }
    }
    h->nranks = nranks; h->rank = rank;
    return ELLCUDA_OK;
}

int ellcuda_exchange_attach(ell_handle h, int32_t nranks, int32_t rank, const void* handles) {
    if (!h || !handles) { set_error("null argument"); return ELLCUDA_E_ARG; }
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
    return ELLCUDA_OK;
}

int ellcuda_set_stream(ell_handle h, void* stream) {
    if (!h) { set_error("null handle"); return ELLCUDA_E_ARG; }
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
    ELL_DEVICE_SCOPE(h->device);
    ELL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return ELLCUDA_OK;
}

int ellcuda_set_timing(ell_handle h, int32_t on) {
    if (!h) { set_error("null handle"); return ELLCUDA_E_ARG; }
    h->timing = on != 0;
    h->timed = false;
    return ELLCUDA_OK;
}

float ellcuda_last_sweep_ms(ell_handle h) {
    if (!h || !h->timed) return -1.0f;
    float ms = -1.0f;
    if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.0f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.0f;
    return ms;
}

int64_t ellcuda_handle_launch_count(ell_handle h) { return h ? h->launches : 0; }

}  // extern "C"
