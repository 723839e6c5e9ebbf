// stream_sym.cu — large-n `Ell` with SYMMETRIC STORAGE (ell_create_desc.storage = ELL_STORAGE_SYMMETRIC): only the lower
// block-triangle of M is kept in HBM, as 64x64 tiles, so one update reads and writes 8·n² bytes instead of 16·n².
//
// The reference stores the full mirrored matrix (src/ell.rs:117-128) and streams it three times per update; the full-storage
// path (stream_update.cu) streams it once each way and is bit-identical.  Here a tile (I,J), I > J, serves BOTH triangles:
//     grad_t[cols of J] += sum over its rows   M[i][c]·g[i]      (column sums, lane = two columns, as in stream_update.cu)
//     grad_t[rows of I] += sum over its cols   M[i][c]·g[c]      (row sums, second pass over the tile in shared memory)
// which regroups the reference's left-to-right row sums; the partial sums are written to fixed slots and added in a fixed
// order by `sym_combine_kernel`, so results are deterministic, identical run to run, and within north_star's 1e-9 of the
// reference (tests hold them to that), not bit-identical.  The deferred rank-1 of the previous cut is applied to the tile in
// the same pass (src/ell.rs:117-128: element (i,j), i >= j, -= (ratio·g̃[i])·g̃[j]), diagonal tiles keep both triangles.
//
// Work split: one CTA of two warps per (pair of column blocks, row segment): column block J has T−J tiles, so J is paired with T−1−J
// (T+1 tiles per pair) and the pair's tile list is cut into `nseg` contiguous chunks.  Tiles are contiguous 32 KB blocks:
// 1-D bulk copies (cp.async.bulk) through a 3-stage mbarrier ring, updated in place, stored back by bulk copy.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "stream_common.cuh"

namespace ellcuda {
namespace {

constexpr int kTile = 64;
constexpr int kTileDoubles = kTile * kTile;
constexpr int kTileBytes = kTileDoubles * 8;
constexpr int kSymStages = 3;
constexpr int kVecBytes = kTile * 8;    // g rows of a tile
constexpr int kPendBytes = kTile * 16;  // {grad_t, ratio·grad_t} rows of a tile

__host__ __device__ inline long long tile_index(long long I, long long J) { return I * (I + 1) / 2 + J; }   // I >= J

struct SymArgs {
    double* Mp;
    const double* g;       // readable for T·64 rows
    const double2* pend;   // {grad_t[i], ratio·grad_t[i]} of the deferred rank-1, padded with zeros
    const Scal* sc;
    double* colpart;       // nseg × npad
    double* rowpart;       // T × npad
    long long npad;
    int T, nseg;
    int exchange;          // 0: flush only (apply the deferred rank-1, produce nothing)
    int use_map;           // 1: tiles move by cp.async.bulk.tensor.2d over the packed buffer seen as a (tiles·64) x 64 matrix
    int rank, count;       // sharded: this launch covers the pairs rank, rank + count, ... (whole: 0, 1)
};


struct SymRing {
    unsigned char *tiles, *pends, *gss;
    double* gcol;
    unsigned long long* bars;
    __device__ explicit SymRing(unsigned char* raw) {
        tiles = raw;
        pends = raw + kSymStages * kTileBytes;
        gss = pends + kSymStages * kPendBytes;
        gcol = reinterpret_cast<double*>(gss + kSymStages * kVecBytes);
        bars = reinterpret_cast<unsigned long long*>(gss + kSymStages * kVecBytes + kVecBytes);
    }
    __device__ __forceinline__ void issue_m(const SymArgs& a, const CUtensorMap* map, int t, long long I, long long J) const {
        const int s = t % kSymStages;
        mbar_expect_tx(&bars[s], kTileBytes);
        if (a.use_map) tma_load_2d(tiles + s * kTileBytes, map, 0, (int)(tile_index(I, J) * kTile), &bars[s]);
        else bulk_load_1d(tiles + s * kTileBytes, a.Mp + tile_index(I, J) * kTileDoubles, kTileBytes, &bars[s]);
    }
    template <bool PENDING>
    __device__ __forceinline__ void issue_v(const SymArgs& a, int t, long long I) const {
        const int s = t % kSymStages;
        mbar_expect_tx(&bars[s], kVecBytes + (PENDING ? kPendBytes : 0));
        bulk_load_1d(gss + s * kVecBytes, a.g + I * kTile, kVecBytes, &bars[s]);
        if (PENDING) bulk_load_1d(pends + s * kPendBytes, a.pend + I * kTile, kPendBytes, &bars[s]);
    }
};
constexpr size_t kSymSmem = (size_t)kSymStages * (kTileBytes + kPendBytes + kVecBytes) + kVecBytes + kSymStages * 8 + 128;

struct Chunk {   // this warp's part of its pair's tile list
    int Ja, Jb, lenA, v0, ntiles;
    __device__ __forceinline__ void tile_of(int t, int& I, int& J) const {
        const int v = v0 + t;
        if (v < lenA) { J = Ja; I = Ja + v; }
        else { J = Jb; I = Jb + (v - lenA); }
    }
};

// Two warps per CTA: in pass 1 warp w owns columns 32w..32w+31 of the tile (one per lane, rows ascending), in pass 2 rows
// 32w..32w+31 (one per lane).  With a single warp the tile cost 1730 dependent instructions at ~3.5 clocks each = 3 us against
// 2.7 us of HBM time per tile (ncu: 28 % issue active, 71 % of the copy bandwidth); split over two warps it is well below.
template <bool PENDING, bool SCALE>
__device__ __forceinline__ void sym_body(const SymArgs& a, const CUtensorMap* map, const SymRing& ring, const Chunk& ch, int pre, double scale) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int lcol = 32 * wid + lane;   // this thread's column in pass 1, row in pass 2
    const int seg = blockIdx.y;
    if (threadIdx.x == 0)
        for (int t = 0; t < pre; ++t) {
            int I, J;
            ch.tile_of(t, I, J);
            ring.issue_v<PENDING>(a, t, I);   // the tile itself was requested in the prologue
        }
    double* cp = a.colpart + (long long)seg * a.npad;
    if (a.exchange) {   // a slot the chunk does not reach must still read as zero in the combine
        cp[(long long)ch.Ja * kTile + lcol] = 0.0;
        cp[(long long)ch.Jb * kTile + lcol] = 0.0;
    }
    int curJ = -1;
    double acc = 0.0, unused = 0.0, gt_c = 0.0, rg_c = 0.0;
    for (int t = 0; t < ch.ntiles; ++t) {
        int I, J;
        ch.tile_of(t, I, J);
        if (J != curJ) {   // a new column block (at most once per chunk after the first tile)
            if (curJ >= 0 && a.exchange) cp[(long long)curJ * kTile + lcol] = acc;
            curJ = J;
            acc = 0.0;
            const long long c = (long long)J * kTile + lcol;
            if (PENDING) {
                const double2 p0 = a.pend[c];
                gt_c = p0.x; rg_c = p0.y;
            }
            __syncthreads();
            ring.gcol[lcol] = a.g[c];   // g of the columns (row sums)
            __syncthreads();
        }
        const int s = t % kSymStages;
        mbar_wait(&ring.bars[s], (unsigned)((t / kSymStages) & 1));
        double* tile = reinterpret_cast<double*>(ring.tiles + s * kTileBytes);
        const double2* pd = reinterpret_cast<const double2*>(ring.pends + s * kPendBytes);
        const double* gs = reinterpret_cast<const double*>(ring.gss + s * kVecBytes);
        // pass 1: the deferred rank-1 in place, and the column sums (rows ascending)
        if (I == J) tile_rows<kTile, 2, PENDING, SCALE, kTile, 1>(tile + 32 * wid, gs, pd, lane, kTile, 0, lcol, gt_c, 0.0, rg_c, 0.0, scale, acc, unused);
        else tile_rows<kTile, 1, PENDING, SCALE, kTile, 1>(tile + 32 * wid, gs, pd, lane, kTile, 0, 0, gt_c, 0.0, rg_c, 0.0, scale, acc, unused);
        if (I != J && a.exchange) {
            // pass 2: the tile also stands for M[c][i] of the upper triangle: row sums against g of the column block.  Thread q takes
            // row q and visits the columns in an order that keeps a half-warp on 16 different banks at every step.
            if (PENDING) __syncthreads();
            // column order per thread: 16q + (k ^ lane%16) — a fixed permutation (deterministic), two accumulation chains,
            // and only 16 index computations per tile (the rotation (c + lane) % 64 cost 11 instructions per element)
            double ra = 0.0, rb = 0.0;
            const double* row = tile + lcol * kTile;
            const int l15 = lane & 15;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int kk = k ^ l15;
                const double* rp = row + kk;
                const double* gp = ring.gcol + kk;
                ra += rp[0] * gp[0];
                rb += rp[16] * gp[16];
                ra += rp[32] * gp[32];
                rb += rp[48] * gp[48];
            }
            a.rowpart[(long long)J * a.npad + (long long)I * kTile + lcol] = ra + rb;
        }
        if (PENDING) fence_async_smem();  // generic-proxy writes -> visible to the bulk store
        __syncthreads();
        if (threadIdx.x == 0) {
            if (PENDING) {
                if (a.use_map) tma_store_2d(map, 0, (int)(tile_index(I, J) * kTile), tile);
                else bulk_store_1d(a.Mp + tile_index(I, J) * kTileDoubles, tile, kTileBytes);
                bulk_commit();
            }
            const int nt = t + kSymStages - 1;  // refill the stage whose store was issued one iteration ago
            if (nt < ch.ntiles) {
                if (PENDING) bulk_wait_read<1>();
                int In, Jn;
                ch.tile_of(nt, In, Jn);
                ring.issue_m(a, map, nt, In, Jn);
                ring.issue_v<PENDING>(a, nt, In);
            }
        }
    }
    if (curJ >= 0 && a.exchange) cp[(long long)curJ * kTile + lcol] = acc;
    if (PENDING && threadIdx.x == 0) bulk_wait_all();
}

// grid = (pairs of column blocks, row segments); block = 64 threads
__global__ void __launch_bounds__(64) sym_sweep_kernel(const __grid_constant__ CUtensorMap map, const SymArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const SymRing ring(smem_raw);
    Chunk ch;
    ch.Ja = a.rank + blockIdx.x * a.count;
    ch.Jb = a.T - 1 - ch.Ja;
    ch.lenA = a.T - ch.Ja;
    const int lenB = ch.Jb != ch.Ja ? a.T - ch.Jb : 0;
    const int total = ch.lenA + lenB;
    const int per = (total + a.nseg - 1) / a.nseg;
    ch.v0 = blockIdx.y * per;
    ch.ntiles = total - ch.v0 < per ? total - ch.v0 : per;
    if (ch.ntiles < 0) ch.ntiles = 0;
    const int pre = ch.ntiles < (kSymStages - 1) ? ch.ntiles : (kSymStages - 1);
    // prologue (may overlap the preceding finish kernel, see stream_update.cu): barriers and the first tiles of M
    if (threadIdx.x == 0) {
        for (int s = 0; s < kSymStages; ++s) mbar_init(&ring.bars[s], 2);
        mbar_fence_init();
        for (int t = 0; t < pre; ++t) {
            int I, J;
            ch.tile_of(t, I, J);
            ring.issue_m(a, &map, t, I, J);
        }
    }
    __syncthreads();
    if (a.exchange) pdl_trigger();
    pdl_wait();
    const int pending = a.sc->pending;
    const double scale = a.sc->scale;
    if (!pending) sym_body<false, false>(a, &map, ring, ch, pre, 1.0);
    else if (scale == 1.0) sym_body<true, false>(a, &map, ring, ch, pre, 1.0);
    else sym_body<true, true>(a, &map, ring, ch, pre, scale);
}

// grad_t[i] = (column sums of block B = i/64: segment slots in order) + (row sums rowpart[J][i], J = 0..B-1), in a FIXED order:
// a CTA of 8 warps handles 32 consecutive elements; warp w adds its eighth of the J range in order (coalesced 256-byte loads,
// several in flight), warp 0 then adds the segment slots and the eight parts in order.  (One thread per element walking all
// B <= 255 row blocks took 41 us at n = 16384: a chain of dependent L2 round trips.)
constexpr int kCombThreads = 256;
__device__ __forceinline__ double sym_partial(const double* __restrict__ colpart, const double* __restrict__ rowpart, long long n,
                                              long long npad, int nseg, double* parts /* shared, 8 x 32 */, int T = 0, int rank = 0,
                                              int count = 1) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long i = (long long)blockIdx.x * 32 + lane;
    const int B = (int)(((long long)blockIdx.x * 32) >> 6);   // the same row block for all 32 elements
    const int per = (B + 7) / 8;
    const int j0 = w * per, j1 = j0 + per < B ? j0 + per : B;
    double acc = 0.0;
    if (i < n) {
#pragma unroll 8
        for (int J = j0; J < j1; ++J) {
            if (count > 1) {   // sharded: row blocks of column blocks this rank does not own hold exact zeros — skip them
                const int pr = J < T - 1 - J ? J : T - 1 - J;
                if (pr % count != rank) continue;
            }
            acc += __ldcg(rowpart + (long long)J * npad + i);
        }
    }
    parts[w * 32 + lane] = acc;
    __syncthreads();
    double s = 0.0;
    if (w == 0 && i < n) {
#pragma unroll 8
        for (int k = 0; k < nseg; ++k) s += __ldcg(colpart + (long long)k * npad + i);   // unrolled: the loads go out together, the adds stay in order
#pragma unroll
        for (int q = 0; q < 8; ++q) s += parts[q * 32 + lane];
    }
    return s;   // valid in warp 0
}

__global__ void __launch_bounds__(kCombThreads) sym_combine_kernel(const double* __restrict__ colpart, const double* __restrict__ rowpart,
                                                                   double* __restrict__ out, long long n, long long npad, int nseg) {
    ELL_STATIC_SHARED double parts[8 * 32];
    pdl_trigger();
    pdl_wait();
    const double s = sym_partial(colpart, rowpart, n, npad, nseg, parts);
    const long long i = (long long)blockIdx.x * 32 + (threadIdx.x & 31);
    if (threadIdx.x < 32 && i < n) out[i] = s;
}

// Sharded symmetric storage: every rank has partial sums for ALL n elements (from its tiles), so the exchange is a reduction.
// Phase 1 (this kernel): element i's local partial goes, as a flagged line, to the rank that reduces i (slices of `slice`).
// FUSED (the peers are other GPUs, attached by IPC): the thread whose element lies in this rank's slice goes straight on to
// phase 2 — it depends only on its own store and on the other GPUs' kernels, never on a later launch of this GPU.
__device__ __forceinline__ void sym_reduce_one(long long n, const PeerTable& pt, Scal* sc, long long slice, long long j, unsigned long long epoch);

template <bool FUSED>
__global__ void __launch_bounds__(kCombThreads) sym_combine_push_kernel(const double* __restrict__ colpart, const double* __restrict__ rowpart,
                                                                        long long n, long long npad, int nseg, PeerTable pt, Scal* sc,
                                                                        long long slice) {
    ELL_STATIC_SHARED double parts[8 * 32];
    pdl_trigger();
    pdl_wait();
    const double s = sym_partial(colpart, rowpart, n, npad, nseg, parts, (int)(npad >> 6), pt.rank, pt.nranks);
    const long long i = (long long)blockIdx.x * 32 + (threadIdx.x & 31);
    if (threadIdx.x >= 32 || i >= n || sc->failed) return;
    const unsigned long long epoch = sc->epoch;
    const int o = (int)(i / slice);
    ll_store(xblk_parts(pt.blk[o], n, (int)(epoch & 1ULL), pt.nranks, pt.rank, slice) + (i - (long long)o * slice), s, (unsigned)epoch);
    if (FUSED && o == pt.rank) sym_reduce_one(n, pt, sc, slice, i - (long long)o * slice, epoch);
}

// Phase 2: this rank's slice — the P partials in rank order (identical on every rank by construction: one rank does it),
// then the sum as a flagged line into every rank's grad_t lines, which finish_kernel polls as for a gathered exchange.
__device__ __forceinline__ void sym_reduce_one(long long n, const PeerTable& pt, Scal* sc, long long slice, long long j, unsigned long long epoch) {
    const long long i = (long long)pt.rank * slice + j;
    const int par = (int)(epoch & 1ULL);
    const unsigned flag = (unsigned)epoch;
    // all P lines are requested before any is checked (one round trip instead of P); a line that has not arrived yet is polled
    unsigned wds[kMaxRanks][4];
#pragma unroll
    for (int r = 0; r < kMaxRanks; ++r)
        if (r < pt.nranks) {
            const uint4* line = xblk_parts(pt.blk[pt.rank], n, par, pt.nranks, r, slice) + j;
            ll_peek(line, wds[r]);
        }
    double t = 0.0;
#pragma unroll
    for (int r = 0; r < kMaxRanks; ++r)
        if (r < pt.nranks) {
            double v = __hiloint2double((int)wds[r][2], (int)wds[r][0]);
            if (wds[r][1] != flag || wds[r][3] != flag) {
                if (!ll_load(xblk_parts(pt.blk[pt.rank], n, par, pt.nranks, r, slice) + j, flag, v)) {
                    atomicExch(&sc->failed, 1);   // a peer never delivered: the handle is dead (finish reports Unknown from now on)
                    return;
                }
            }
            t += v;   // rank order
        }
    for (int r = 0; r < pt.nranks; ++r) ll_store(xblk_lines(pt.blk[r], n, par) + i, t, flag);
}

__global__ void __launch_bounds__(256) sym_reduce_kernel(long long n, PeerTable pt, Scal* sc, long long slice) {
    pdl_trigger();
    pdl_wait();
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= slice || (long long)pt.rank * slice + j >= n || sc->failed) return;
    sym_reduce_one(n, pt, sc, slice, j, sc->epoch);
}

__global__ void sym_fill_diag_kernel(double* __restrict__ Mp, long long n, const double* __restrict__ diag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long I = i >> 6, r = i & 63;
    Mp[tile_index(I, I) * kTileDoubles + r * kTile + r] = diag ? diag[i] : 1.0;
}

__global__ void sym_unpack_kernel(const double* __restrict__ Mp, double* __restrict__ out, long long n) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    for (long long i = blockIdx.y; i < n; i += gridDim.y) {
        const long long I = i >> 6, J = j >> 6, r = i & 63, c = j & 63;
        out[i * n + j] = I >= J ? Mp[tile_index(I, J) * kTileDoubles + r * kTile + c] : Mp[tile_index(J, I) * kTileDoubles + c * kTile + r];
    }
}

}  // namespace

#ifdef __CUDACC__   // host side: left out when tests/cpp/emu_sym.cpp compiles the kernels of this file for the host
int sym_alloc(ellcuda_ctx* h) {
    h->T = (h->n + kTile - 1) / kTile;
    const long long npad = h->T * kTile;
    const long long all_pairs = (h->T + 1) / 2;
    long long npairs = (all_pairs - h->sym_rank + h->sym_count - 1) / h->sym_count;          // the pairs this handle owns
    if (npairs < 1) npairs = 1;
    h->sym_slice = (h->n + h->sym_count - 1) / h->sym_count;
    int nseg = h->reduce_mode > 0 ? h->reduce_mode : (int)((256 + npairs - 1) / npairs);   // ~256 streaming CTAs
    if (nseg < 1) nseg = 1;
    if (nseg > 16) nseg = 16;   // measured on the 8-way emulation (tools/profile_symslab.py): 8 -> 119 us per rank, 12 -> 93, 16 -> 82.5, 24 -> 100, 32 -> 92
    h->sym_nseg = nseg;
    const size_t tiles = (size_t)tile_index(h->T - 1, h->T - 1) + 1;
    ELL_CUDA_TRY(cudaMalloc(&h->Mp, tiles * kTileBytes));
    ELL_CUDA_TRY(cudaMalloc(&h->colpart, (size_t)nseg * npad * 8));
    ELL_CUDA_TRY(cudaMalloc(&h->rowpart, (size_t)h->T * npad * 8));
    ELL_CUDA_TRY(cudaMemset(h->colpart, 0, (size_t)nseg * npad * 8));
    ELL_CUDA_TRY(cudaMemset(h->rowpart, 0, (size_t)h->T * npad * 8));
    // the packed tiles as a tall matrix of 64 columns: tile k = rows [64k, 64k+64), so a tile is one 64x64 TMA box
    return make_tensor_map(&h->map, h->Mp, (long long)tiles * kTile, kTile, kTile, kTile, kTile);
}

void sym_free(ellcuda_ctx* h) {
    cudaFree(h->Mp); cudaFree(h->colpart); cudaFree(h->rowpart);
    h->Mp = h->colpart = h->rowpart = nullptr;
}

int sym_upload(ellcuda_ctx* h, const double* mq, const double* diag) {
    cudaStream_t s = h->stream;
    const long long n = h->n, T = h->T;
    const size_t tiles = (size_t)tile_index(T - 1, T - 1) + 1;
    ELL_CUDA_TRY(cudaMemsetAsync(h->Mp, 0, tiles * kTileBytes, s));
    if (!mq) {
        double* ddiag = nullptr;
        if (diag) {
            ELL_CUDA_TRY(cudaMallocAsync((void**)&ddiag, (size_t)n * 8, s));
            ELL_CUDA_TRY(cudaMemcpyAsync(ddiag, diag, (size_t)n * 8, cudaMemcpyHostToDevice, s));
        }
        sym_fill_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h->Mp, n, ddiag);
        if (ddiag) cudaFreeAsync(ddiag, s);
        ELL_CUDA_TRY(cudaGetLastError());
        ELL_CUDA_TRY(cudaStreamSynchronize(s));
        return ELLCUDA_OK;
    }
    // one tile row at a time: pack (I,0..I) on the host, copy
    std::vector<double> stage((size_t)T * kTileDoubles);
    for (long long I = 0; I < T; ++I) {
        std::fill(stage.begin(), stage.begin() + (size_t)(I + 1) * kTileDoubles, 0.0);
        for (long long J = 0; J <= I; ++J)
            for (long long r = 0; r < kTile; ++r) {
                const long long i = I * kTile + r;
                if (i >= n) break;
                const long long cmax = std::min<long long>(kTile, n - J * kTile);
                memcpy(&stage[(size_t)J * kTileDoubles + r * kTile], mq + i * n + J * kTile, (size_t)cmax * 8);
            }
        ELL_CUDA_TRY(cudaMemcpyAsync(h->Mp + tile_index(I, 0) * kTileDoubles, stage.data(), (size_t)(I + 1) * kTileBytes, cudaMemcpyHostToDevice, s));
        ELL_CUDA_TRY(cudaStreamSynchronize(s));
    }
    return ELLCUDA_OK;
}

int sym_sweep(ellcuda_ctx* h, const double* g_dev, int exchange, double* plain_out) {
    cudaStream_t s = h->stream;
    static std::mutex mu;
    static unsigned long long done_mask = 0;
    {
        std::lock_guard<std::mutex> lk(mu);   // per-device opt-in to > 48 KB of dynamic shared memory
        if (!(h->device < 64 && ((done_mask >> h->device) & 1ULL))) {
            ELL_CUDA_TRY(cudaFuncSetAttribute(sym_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSymSmem));
            if (h->device < 64) done_mask |= 1ULL << h->device;
        }
    }
    SymArgs a;
    a.Mp = h->Mp; a.g = g_dev; a.pend = h->pend; a.sc = h->sc; a.colpart = h->colpart; a.rowpart = h->rowpart;
    a.npad = h->T * kTile; a.T = (int)h->T; a.nseg = h->sym_nseg; a.exchange = exchange;
    a.use_map = ELL_ENV_INT("ELLCUDA_SYM_TMA", 1);   // 0: plain 1-D bulk copies of the contiguous 32 KB tile (experiments)
    a.rank = h->sym_rank; a.count = h->sym_count;
    const long long all_pairs = (h->T + 1) / 2;
    const long long own_pairs = all_pairs > h->sym_rank ? (all_pairs - h->sym_rank + h->sym_count - 1) / h->sym_count : 0;
    if (own_pairs == 0 && !exchange) return ELLCUDA_OK;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(own_pairs > 0 ? own_pairs : 1), (unsigned)h->sym_nseg); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = kSymSmem; cfg.stream = s;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    if (own_pairs > 0) {
        ELL_CUDA_TRY(cudaLaunchKernelEx(&cfg, sym_sweep_kernel, h->map, a));
        h->launches += 1;
    }
    if (!exchange) return ELLCUDA_OK;
    if (h->sym_count > 1 && !plain_out) {   // sharded: the local partial of every element goes to the rank that reduces it
        if (h->nranks != h->sym_count) { set_error("a sharded symmetric handle must be attached to its peers before it is updated"); return ELLCUDA_E_ARG; }
        PeerTable pt;
        pt.nranks = h->nranks; pt.rank = h->rank;
        for (int r = 0; r < kMaxRanks; ++r) pt.blk[r] = r < h->nranks ? h->peer_blk[r] : nullptr;
        cudaLaunchConfig_t cc{};
        cc.gridDim = dim3((unsigned)((h->n + 31) / 32)); cc.blockDim = dim3(kCombThreads); cc.dynamicSmemBytes = 0; cc.stream = s;
        cc.attrs = attr; cc.numAttrs = pdl_enabled() ? 1 : 0;
        if (h->sym_fused)
            ELL_CUDA_TRY(cudaLaunchKernelEx(&cc, sym_combine_push_kernel<true>, (const double*)h->colpart, (const double*)h->rowpart, h->n,
                                            a.npad, a.nseg, pt, h->sc, h->sym_slice));
        else
            ELL_CUDA_TRY(cudaLaunchKernelEx(&cc, sym_combine_push_kernel<false>, (const double*)h->colpart, (const double*)h->rowpart, h->n,
                                            a.npad, a.nseg, pt, h->sc, h->sym_slice));
        h->launches += 1;
        return ELLCUDA_OK;
    }
    cudaLaunchConfig_t cc{};
    cc.gridDim = dim3((unsigned)((h->n + 31) / 32)); cc.blockDim = dim3(kCombThreads); cc.dynamicSmemBytes = 0; cc.stream = s;
    cc.attrs = attr; cc.numAttrs = pdl_enabled() ? 1 : 0;
    ELL_CUDA_TRY(cudaLaunchKernelEx(&cc, sym_combine_kernel, (const double*)h->colpart, (const double*)h->rowpart,
                                    plain_out ? plain_out : h->xblk, h->n, a.npad, a.nseg));
    h->launches += 1;
    return ELLCUDA_OK;
}

int sym_reduce(ellcuda_ctx* h) {
    if (h->sym_count <= 1 || h->sym_fused) return ELLCUDA_OK;   // fused: the combine kernel already did it
    PeerTable pt;
    pt.nranks = h->nranks; pt.rank = h->rank;
    for (int r = 0; r < kMaxRanks; ++r) pt.blk[r] = r < h->nranks ? h->peer_blk[r] : nullptr;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cc{};
    cc.gridDim = dim3((unsigned)((h->sym_slice + 255) / 256)); cc.blockDim = dim3(256); cc.dynamicSmemBytes = 0; cc.stream = h->stream;
    cc.attrs = attr; cc.numAttrs = pdl_enabled() ? 1 : 0;
    ELL_CUDA_TRY(cudaLaunchKernelEx(&cc, sym_reduce_kernel, h->n, pt, h->sc, h->sym_slice));
    h->launches += 1;
    return ELLCUDA_OK;
}

int sym_unpack(ellcuda_ctx* h, double* out_dev) {
    dim3 grid((unsigned)((h->n + 255) / 256), (unsigned)std::min<long long>(h->n, 32768));
    sym_unpack_kernel<<<grid, 256, 0, h->stream>>>(h->Mp, out_dev, h->n);
    ELL_CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return ELLCUDA_OK;
}

int sym_copy(ellcuda_ctx* dst, const ellcuda_ctx* src, cudaStream_t s) {
    const size_t tiles = (size_t)tile_index(src->T - 1, src->T - 1) + 1;
    ELL_CUDA_TRY(cudaMemcpyAsync(dst->Mp, src->Mp, tiles * kTileBytes, cudaMemcpyDeviceToDevice, s));
    return ELLCUDA_OK;
}

#endif  // __CUDACC__

}  // namespace ellcuda
