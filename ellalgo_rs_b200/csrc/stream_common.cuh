// stream_common.cuh — what the large-n translation units (stream_update.cu: full storage, row-shardable; stream_sym.cu:
// symmetric block-triangle storage) share: PTX helpers (mbarrier, TMA / bulk copies, programmatic dependent launch), the
// device-resident scalar block, the peer-exchange line format, the per-tile update + column-sum loop, and the handle record.
#pragma once
#include <cuda.h>

#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "stream_ptx.cuh"

namespace ellcuda {
namespace {

constexpr int kColsMax = 64;  // columns per warp: 64 (2 per lane), 32 (1 per lane) or 16 (lanes 0..15) — narrower blocks give
                              // bit-identical mode enough warps at medium n (n/COLS column blocks)
constexpr int kRowsDefault = 64;    // rows per TMA tile   (measured at n=16384: 64x3 682 us, 32x6 693, 48x4 692,
constexpr int kStagesDefault = 3;   // smem ring depth      32x4 709, 48x3 753, 32x3 840, 16x8 831, 64x4 1219 [1 CTA/SM])
constexpr int kRowsMax = 64;
// per tile: COLS*ROWS*8 bytes of M, ROWS*8 bytes of g and ROWS*16 bytes of {gt[i], ratio*gt[i]}

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
    int failed;     // sticky: a peer wait timed out in some finish CTA; every later finish reports Unknown and the sweeps stop
    int _pad;       // publishing, so that the peers time out too instead of consuming a half-updated state
};

// (the inline PTX — mbarrier, TMA / bulk copies, flagged lines, programmatic dependent launch — is in stream_ptx.cuh)

// ------------------------------------------------------------------------------------------------ kernels
// Peer exchange (row-sharded handles attached with ellcuda_exchange_attach): the sweep CTA that completes a block of
// 64 columns of grad_t stores it straight into EVERY rank's exchange block over NVLink (peer pointers) as flagged
// 16-byte lines {lo32, epoch, hi32, epoch} — each 8-byte half carries its own flag and 8-byte stores are never torn,
// so a reader that sees both flags equal to the update's epoch has the value (the "LL" scheme of collective
// libraries).  There is no fence, no counter and no separate flag word on the path: one NVLink store latency after a
// column block is summed, every rank can use it.  `finish_kernel` on every rank polls the n lines.  An exchange block
// holds two line buffers (epoch parity) so that a rank one update ahead never overwrites data a slower rank still
// reads.  No collective call, no host involvement.
constexpr int kMaxRanks = 16;
struct PeerTable {
    double* blk[kMaxRanks];   // rank r's exchange block (layout: xblk_* helpers below)
    int nranks;               // 0: no peer exchange (grad_t goes to this rank's buffer 0 only)
    int rank;
};
// exchange block, in doubles: grad_t[n] plain (handles without peer exchange; ellcuda_grad_t_ptr) | pad to 16 bytes |
// lines[2][n] of 16 bytes (peer exchange, buffer = epoch parity)
__host__ __device__ inline size_t xblk_plain_doubles(long long n) { return (size_t)((n + 1) & ~1LL); }
__host__ __device__ inline size_t xblk_doubles(long long n) { return xblk_plain_doubles(n) + (size_t)4 * n; }
// sharded symmetric storage adds, after the lines: parts[2][nranks][slice] lines — rank r's partial sums for the elements this
// rank reduces (buffer = epoch parity)
__host__ __device__ inline uint4* xblk_parts(double* blk, long long n, int parity, int nranks, int src, long long slice) {
    return reinterpret_cast<uint4*>(blk + xblk_doubles(n)) + ((long long)parity * nranks + src) * slice;
}
__host__ __device__ inline uint4* xblk_lines(double* blk, long long n, int parity) {
    return reinterpret_cast<uint4*>(blk + xblk_plain_doubles(n)) + (long long)parity * n;
}
// MODE 0: every row of the tile is above the block's columns (i < c):   update = (ratio·gt[c]) · gt[i]
// MODE 1: every row is at or below them (i >= c):                        update = (ratio·gt[i]) · gt[c]
// MODE 2: the tile crosses the diagonal: choose per element.             (src/ell.rs:118-127: hi = max, lo = min)
// STRIDE = doubles per tile row, PERLANE = columns per lane: a 64-wide tile is walked either by one warp with two columns per
// lane (16-byte accesses) or by two warps with one column per lane each (the second warp's base pointer is 32 columns in).
template <int COLS, int MODE, bool PENDING, bool SCALE, int STRIDE = COLS, int PERLANE = (COLS == 64 ? 2 : 1)>
__device__ __forceinline__ void tile_rows(double* __restrict__ tile_raw, const double* __restrict__ gs, const double2* __restrict__ pd,
                                          int lane, int nrows, int row0, int col0, double gt_c0, double gt_c1, double rg_c0,
                                          double rg_c1, double scale, double& acc0, double& acc1) {
    constexpr bool TWO = PERLANE == 2;                   // two columns per lane (16-byte accesses) or one
    if (COLS < 32 && lane >= COLS) return;               // 16-column blocks: the upper half-warp has no column
#pragma unroll 8
    for (int r = 0; r < nrows; ++r) {  // nrows < tile height only in the last tile of a segment (rows past n are not M)
        double2 m;
        if (TWO) m = reinterpret_cast<double2*>(tile_raw)[r * (STRIDE / 2) + lane];
        else { m.x = tile_raw[r * STRIDE + lane]; m.y = 0.0; }
        const double gi = gs[r];       // broadcast
        if (PENDING) {
            const double2 p = pd[r];   // broadcast: {gt[i], ratio*gt[i]}
            double u0, u1 = 0.0;
            if (MODE == 0) {
                u0 = rg_c0 * p.x;
                if (TWO) u1 = rg_c1 * p.x;
            } else if (MODE == 1) {
                u0 = p.y * gt_c0;
                if (TWO) u1 = p.y * gt_c1;
            } else {
                const int i = row0 + r;
                u0 = (i >= col0) ? p.y * gt_c0 : rg_c0 * p.x;
                if (TWO) u1 = (i >= col0 + 1) ? p.y * gt_c1 : rg_c1 * p.x;
            }
            m.x -= u0;
            if (TWO) m.y -= u1;
            if (SCALE) {
                m.x *= scale;  // mq *= kappa   (src/ell.rs:132-135)
                if (TWO) m.y *= scale;
            }
            if (TWO) reinterpret_cast<double2*>(tile_raw)[r * (STRIDE / 2) + lane] = m;
            else tile_raw[r * STRIDE + lane] = m.x;
        }
        acc0 += m.x * gi;  // grad_t[c] += M[i][c] * g[i], i ascending
        if (TWO) acc1 += m.y * gi;
    }
}

// ------------------------------------------------------------------------------------------------ tensor map
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline int make_tensor_map(CUtensorMap* map, void* base, long long rows, long long cols, long long ld, int box_rows, int box_cols) {
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
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
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

// ELLCUDA_PDL=0 switches programmatic dependent launch off (plain stream order between the kernels of an update)
inline bool pdl_enabled() { return ELL_ENV_INT("ELLCUDA_PDL", 1) != 0; }

}  // namespace
}  // namespace ellcuda

// =================================================================================================== handle
struct ellcuda_ctx {
    int kind = ELL_KIND_ELL;    // ELL_KIND_ELLSTABLE handles (n <= 2048) use only the `st_*` members, xc, gstage, status_host
    double* st_qu = nullptr;    // EllStable: packed upper triangle incl. diagonal, n(n+1)/2
    double* st_scal = nullptr;  // EllStable: {kappa, tsq}
    int* st_status = nullptr;   // EllStable: CutStatus of the last update
    int device = 0;
    long long n = 0, w = 0, c0 = 0, ld = 0, npad = 0, ld_part = 0, ncb = 0;
    int nseg = 1, reduce_mode = 0, no_defer = 0, use_parallel = 1;
    int tile_rows = ellcuda::kRowsDefault, stages = ellcuda::kStagesDefault;   // sweep tile shape (ELLCUDA_SWEEP=rows,stages)
    int cols = ellcuda::kColsMax;   // columns per sweep warp: 64, 32 or 16 (chosen so that enough warps stream)
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    double* S = nullptr;        // n × ld, column block [c0, c0+w) of the symmetric matrix
    double* xc = nullptr;       // n
    double* gpad = nullptr;     // npad: padded copy of the gradient (when n is not a multiple of the tile height), zero tail
    double* gzero = nullptr;    // npad zeros (flush sweeps)
    double* gstage = nullptr;   // n: landing buffer for host gradients (ellcuda_update)
    const double* gsrc = nullptr;   // gradient of the update in flight (read again by finish)
    int* status_host = nullptr; // pinned, CutStatus read-back
    double* xblk = nullptr;     // exchange block (IPC-shareable cudaMalloc): grad_t[2][n], flags, omega parts
    double2* pend = nullptr;    // npad + kColsMax + 2: deferred rank-1 vectors {grad_t, ratio·grad_t}
    double* partial = nullptr;  // nseg × ld_part
    unsigned* counters = nullptr;   // ncb (per column block) + 1 (blocks done) + 1 (finish CTAs arrived)
    ellcuda::Scal* sc = nullptr;
    int nranks = 0, rank = 0;                    // > 0 once ellcuda_exchange_attach* succeeded
    double* peer_blk[ellcuda::kMaxRanks] = {};   // every rank's exchange block as mapped in this process
    bool peer_opened[ellcuda::kMaxRanks] = {};   // mapped by cudaIpcOpenMemHandle (must be closed)
    CUtensorMap map;
    // Ell::new_with_matrix with a NON-symmetric matrix (src/ell.rs:31-41 takes any Arr): until the first successful cut the
    // reference multiplies by the rows as given (src/ell.rs:102) and leaves the matrix alone; the first successful cut then
    // rewrites the upper triangle from the (updated) lower one (:117-128).  `S` holds the mirrored lower triangle from the
    // start (what every rank-1 is applied to); `A` holds the matrix as given, transposed (A[i][c] = mq[c0+c][i], so that
    // column sums over i are the reference's row sums), and is read by the sweeps while `asym_live`; the first Success
    // (learned by a synchronous status read in finish) retires it.
    bool asym_live = false;
    double* A = nullptr;
    CUtensorMap mapA;
    // ellcuda_create_sharded: this handle is a GROUP — one slab handle per device of this process, peers attached by
    // pointer; it owns no device memory itself and every entry point fans out over `parts`
    // ELL_STORAGE_SYMMETRIC (stream_sym.cu): only the lower block-triangle of 64x64 tiles is stored and streamed — half the
    // HBM traffic per update; sums are regrouped (tile row sums + column sums), so results are held to 1e-9, not bit-identical
    bool sym = false;
    double* Mp = nullptr;       // T(T+1)/2 tiles of 64x64 doubles, tile (I,J), I >= J, at index I(I+1)/2 + J; diagonal tiles full
    double* colpart = nullptr;  // sym_nseg x npad64: column sums per row-segment slot
    double* rowpart = nullptr;  // T x npad64: rowpart[J][i] = sum over the columns of block J of M[i][c] g[c], for rows below block J
    long long T = 0;            // tile rows
    int sym_nseg = 1;
    int sym_rank = 0, sym_count = 1;   // sharded symmetric storage: this handle keeps the pairs p = sym_rank (mod sym_count) current
    long long sym_slice = 0;    // ceil(n / sym_count): the elements of grad_t this rank reduces
    bool sym_fused = false;     // peers are other processes / GPUs (IPC attach): the combine kernel also reduces and broadcasts its slice
    std::vector<ellcuda_ctx*> parts;
    double* ytmp = nullptr;     // group members: slab of a dot_mv result before it is copied to the caller's vector
    int which = 0;
    ellcuda::Cut cut{0.0, 0.0, false};
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    long long launches = 0;
};


namespace ellcuda {
// stream_sym.cu — symmetric block-triangle storage
int sym_alloc(ellcuda_ctx* h);
void sym_free(ellcuda_ctx* h);
int sym_upload(ellcuda_ctx* h, const double* mq_host, const double* diag_host);          // synchronous
int sym_sweep(ellcuda_ctx* h, const double* g_dev, int exchange, double* plain_out);     // sweep (+ combine into grad_t / plain_out)
int sym_reduce(ellcuda_ctx* h);    // sharded: sum the ranks' partials of this rank's slice, broadcast the result (after sym_sweep on every rank)
int sym_unpack(ellcuda_ctx* h, double* out_dev_n_by_n);                                  // the full matrix, row-major, device
int sym_copy(ellcuda_ctx* dst, const ellcuda_ctx* src, cudaStream_t s);
}  // namespace ellcuda
