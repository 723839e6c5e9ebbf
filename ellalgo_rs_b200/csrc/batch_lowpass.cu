// batch_lowpass.cu — the headline path (LowpassOracle, n = 32, cutting_plane_optim on Ell): one warp per problem,
// persistent, with M held in TENSOR MEMORY instead of registers.
//
// Why: the generic warp-per-problem kernel (batch_warp.cu) runs this workload at 88 % of the shared-memory data pipe
// (profiles/r01c_prof_batch3_*): per iteration it moves ~595 shared-memory wavefronts — 314 for spectrum rows (36 of them
// bank conflicts where a round wraps around the band), 140 for re-reading the broadcast copy of x every 32-row round,
// 141 for the update — while its registers are full (M's row = 64 of 128).  Tensor memory is a separate datapath
// (tools/ubench_tmem.cu: 410 B/clk/SM next to shared memory's 128 B/clk/SM), so:
//   * M lives in TMEM between updates (TMEM lane 32*(warp%4)+i = row i, 64 columns per warp; tcgen05.ld before M.g,
//     tcgen05.st after the rank-1 update);
//   * that frees 64 registers during the oracle phase for x, read once per oracle call instead of once per round;
//   * a round of 32 consecutive rows k0..k0+31 is evaluated with thread L on the row congruent to L mod 32 (traversal
//     position p = (L - k0) mod 32); a round never crosses the band's wrap-around point, so the rows of a half-warp always
//     sit in distinct banks.  "First violated row in traversal order" = lowest set bit of the ballot rotated by k0.
// 470 wavefronts per iteration, 503 M iter/s against 445 (profiles/r01c_prof_batch5_*).
// Experimental variants (ELLCUDA_LOWPASS_VARIANT) additionally read the last 8 columns of every spectrum row from TMEM:
// row r is stored in lane r%32, column block r/32 (TMEM is lane-addressed at a warp-uniform column); a window straddles
// two blocks, both are loaded and each thread keeps its own (variant 1), or rounds are aligned to blocks (variant 2).
// Fewer wavefronts (396) but more instructions or rounds: 490 / 482 M iter/s, so they are not the default.
//
// Arithmetic and its order are those of batch_warp.cu (= the reference's): every row dot is summed left to right
// over j = 0..31, M.g likewise, omega sequentially over k; compiled -fmad=false.
// Reference: src/oracles/lowpass_oracle.rs:58-150, src/cutting_plane.rs:286-313, src/ell.rs:97-137.
#include <cstdlib>

#include "common.cuh"
#include "tmem.cuh"

namespace ellcuda {
namespace {

constexpr int kN = 32;
constexpr int kTabStride = 33;          // doubles per shared-memory row (odd: rows of one half-warp hit distinct banks)
constexpr int kMCols = 64;              // one M row per lane: 32 doubles
constexpr int kTabCol0 = 4 * kMCols;    // 4 warps per lane quarter keep their M in columns [0, 256)
constexpr int kTabBlkCols = 16;         // 32-bit columns per 32-row block of the spectrum tail (8 doubles per row)
constexpr int kScratch = 5 * 32;        // doubles per warp: sx | sg | sgt | sp | srg

// M row <-> two 32-column TMEM transfers
__device__ __forceinline__ void m_load(uint32_t taddr, double (&m)[32]) {
    uint32_t r0[32], r1[32];
    tm_ld32(taddr, r0);
    tm_ld32(taddr + 32, r1);
    tm_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        m[j] = mk_double(r0[2 * j], r0[2 * j + 1]);
        m[16 + j] = mk_double(r1[2 * j], r1[2 * j + 1]);
    }
}
__device__ __forceinline__ void m_store(uint32_t taddr, const double (&m)[32]) {
    uint32_t r0[32], r1[32];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        r0[2 * j] = (uint32_t)__double2loint(m[j]);
        r0[2 * j + 1] = (uint32_t)__double2hiint(m[j]);
        r1[2 * j] = (uint32_t)__double2loint(m[16 + j]);
        r1[2 * j + 1] = (uint32_t)__double2hiint(m[16 + j]);
    }
    tm_st32(taddr, r0);
    tm_st32(taddr + 32, r1);
}

struct Ctx {
    int lane;
    const double* tab;   // shared-memory spectrum, kTabStride doubles per row
    double* ws;          // per-warp scratch
    uint32_t tm_tab;     // TMEM address (this warp's lane quarter) of column block 0 of the spectrum tail
    __device__ __forceinline__ double* sx() const { return ws; }
    __device__ __forceinline__ double* sg() const { return ws + 32; }
    __device__ __forceinline__ double* sgt() const { return ws + 64; }
    __device__ __forceinline__ double* sp() const { return ws + 96; }
    __device__ __forceinline__ double* srg() const { return ws + 128; }
};

__device__ __forceinline__ double seq_sum32(const double* s) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 32; k += 2) {
        const double2 v = *reinterpret_cast<const double2*>(s + k);
        acc += v.x;
        acc += v.y;
    }
    return acc;
}

// kTabTm: trailing columns of each spectrum row read from TMEM (0 or 8).  kAligned: rounds end at multiples of 32 rows, so
// the rows of a round share one TMEM column block (one load, no per-thread select) at the price of more, partly idle rounds.
template <int kTabTm, bool kAligned>
struct Lowpass32 {  // src/oracles/lowpass_oracle.rs:7-150
    int idx1, idx2, idx3, nwpass, nwstop, mdim, kmax;
    double lp_sq, up_sq, sp_sq, fmax;

    __device__ void init(const BatchArgs& a, long long pid) {  // :23-52
        const double* p = a.params + pid * a.params_stride;
        mdim = a.aux[0];
        nwpass = (int)floor(p[0] * (double)(mdim - 1)) + 1;
        nwstop = (int)floor(p[1] * (double)(mdim - 1)) + 1;
        lp_sq = p[2]; up_sq = p[3]; sp_sq = p[4];
        idx1 = -1; idx2 = nwpass - 1; idx3 = nwstop - 1;
        fmax = -INFINITY; kmax = -1;
    }

    // One round-robin band sweep (BAND 0 pass-band :62-78, 1 stop-band :83-104, 2 transition band :106-119) in rounds of
    // up to 32 consecutive rows, stopping at the first row in traversal order that violates.
    template <int BAND>
    __device__ __forceinline__ bool sweep(const Ctx& c, int lo, int hi, int& cursor, double& g, Cut& cut,
                                          const double (&xr)[32]) {
        const int len = hi - lo;
        double best_v = -INFINITY;
        int best_c = 0x7fffffff, best_k = -1;
        int k0 = cursor + 1;
        if (k0 >= hi) k0 = lo;
        for (int done = 0; done < len;) {
            int seg = hi - k0;
            if (seg > 32) seg = 32;
            if (seg > len - done) seg = len - done;
            const int rot = k0 & 31;
            if (kAligned && seg > 32 - rot) seg = 32 - rot;
            const int p = (c.lane - rot) & 31;          // traversal position of this thread's row inside the round
            const int row = k0 + p;                     // congruent to lane mod 32
            const bool active = p < seg;
            const int q = k0 >> 5;
            const bool upper = !kAligned && (row >> 5) != q;
            uint32_t ta[16], tb[16];
            if constexpr (kTabTm > 0) {
                tm_ld16(c.tm_tab + (uint32_t)(q * kTabBlkCols), ta);
                if constexpr (!kAligned) tm_ld16(c.tm_tab + (uint32_t)((q + 1) * kTabBlkCols), tb);   // block 15 is allocated, never selected
            }
            const int srow = row < mdim ? row : row - 32;   // idle threads read a valid row of their own residue
            const double* rp = c.tab + srow * kTabStride;
            double val = 0.0;
#pragma unroll
            for (int j = 0; j < kN - kTabTm; ++j) val += rp[j] * xr[j];
            if constexpr (kTabTm > 0) {
                tm_wait_ld();
#pragma unroll
                for (int j = 0; j < kTabTm; ++j) {
                    const uint32_t wlo = (!kAligned && upper) ? tb[2 * j] : ta[2 * j];
                    const uint32_t whi = (!kAligned && upper) ? tb[2 * j + 1] : ta[2 * j + 1];
                    val += mk_double(wlo, whi) * xr[kN - kTabTm + j];
                }
            }
            bool viol;
            if (BAND == 0) viol = active && (val > up_sq || val < lp_sq);
            else if (BAND == 1) viol = active && (val > sp_sq || val < 0.0);
            else viol = active && (val < 0.0);
            const unsigned mask = __ballot_sync(0xffffffffu, viol);
            if (mask) {
                const unsigned by_pos = __funnelshift_r(mask, mask, rot);   // bit p <- thread (p + rot) mod 32
                const int first = __ffs(by_pos) - 1;
                const int kf = k0 + first;
                const double vf = __shfl_sync(0xffffffffu, val, (first + rot) & 31);
                cursor = kf;
                const double e = c.tab[kf * kTabStride + c.lane];
                if (BAND == 0) {
                    if (vf > up_sq) { g = e; cut = parallel_cut(vf - up_sq, vf - lp_sq); }
                    else { g = -e; cut = parallel_cut(-vf + lp_sq, -vf + up_sq); }
                } else if (BAND == 1) {
                    if (vf > sp_sq) { g = e; cut = parallel_cut(vf - sp_sq, vf); }
                    else { g = -e; cut = parallel_cut(-vf, -vf + sp_sq); }
                    // fmax/kmax of a partial sweep are never read (reset at the next sweep, :83-84)
                } else {
                    g = -e; cut = single_cut(-vf);  // ParallelCut(-val, None)
                }
                return true;
            }
            if (BAND == 1) {  // `if val > fmax` in traversal order = first occurrence of the maximum
                if (active && val > best_v) { best_v = val; best_c = done + p; best_k = row; }
            }
            done += seg;
            k0 += seg;
            if (k0 >= hi) k0 = lo;
        }
        // no violation: the cursor has advanced `len` times and sits on the last row visited
        int last = cursor + len;
        if (last >= hi) last -= len;
        cursor = last;
        if (BAND == 1) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, best_v, off);
                const int oc = __shfl_xor_sync(0xffffffffu, best_c, off);
                const int ok = __shfl_xor_sync(0xffffffffu, best_k, off);
                if (ov > best_v || (ov == best_v && oc < best_c)) { best_v = ov; best_c = oc; best_k = ok; }
            }
            fmax = best_v;
            kmax = best_k;
        }
        return false;
    }

    __device__ __forceinline__ bool assess_optim(const Ctx& c, double& gamma, double& g, Cut& cut) {  // :137-150
        sp_sq = gamma;
        double xr[32];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            const double2 v = *reinterpret_cast<const double2*>(c.sx() + j);
            xr[j] = v.x; xr[j + 1] = v.y;
        }
        if (sweep<0>(c, 0, nwpass, idx1, g, cut, xr)) return false;
        fmax = -INFINITY; kmax = -1;
        if (sweep<1>(c, nwstop, mdim, idx3, g, cut, xr)) return false;
        if (sweep<2>(c, nwpass, nwstop, idx2, g, cut, xr)) return false;
        const double x0 = xr[0];
        if (x0 < 0.0) {  // :123-127
            g = c.lane == 0 ? -1.0 : 0.0;
            cut = single_cut(-x0);
            return false;
        }
        g = c.tab[kmax * kTabStride + c.lane];
        cut = parallel_cut(0.0, fmax);
        gamma = fmax;
        return true;
    }
};

template <int kTabTm, bool kAligned, int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 1) lowpass32_tmem_kernel(const BatchArgs a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ uint32_t tmem_base_s;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int rows = a.aux[0];   // 15 * 32

    for (int idx = threadIdx.x; idx < rows * kN; idx += blockDim.x) {
        const int r = idx >> 5, col = idx & 31;
        smem[r * kTabStride + col] = a.table[idx];
    }
    const uint32_t tm_base = tm_alloc_all(&tmem_base_s, warp);   // all 512 columns; the barrier inside also publishes the table
    const uint32_t tm_lane = tm_base + ((uint32_t)(32 * (warp & 3)) << 16);
    if (kTabTm > 0 && warp < 4) {   // one warp per lane quarter stores the tail columns of row 32*b + lane in column block b
        for (int b = 0; b < rows / 32; ++b) {
            uint32_t r[16];
            const double* src = smem + (32 * b + lane) * kTabStride + (kN - 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                r[2 * j] = (uint32_t)__double2loint(src[j]);
                r[2 * j + 1] = (uint32_t)__double2hiint(src[j]);
            }
            tm_st16(tm_lane + kTabCol0 + b * kTabBlkCols, r);
        }
        tm_wait_st();
    }
    tm_fence_sync();

    Ctx c;
    c.lane = lane;
    c.tab = smem;
    c.ws = smem + ((rows * kTabStride + 1) & ~1) + warp * kScratch;
    c.tm_tab = tm_lane + kTabCol0;
    const uint32_t tm_m = tm_lane + (uint32_t)((warp >> 2) * kMCols);
    const CalcConsts cc = CalcConsts::make(kN, a.use_parallel_cut);

    for (;;) {
        unsigned long long pid_u = 0;
        if (lane == 0) pid_u = atomicAdd(a.queue, 1ULL);
        pid_u = __shfl_sync(0xffffffffu, pid_u, 0);
        const long long pid = (long long)pid_u;
        if (pid >= a.count) break;

        {   // ---- Ell::new / new_with_scalar (src/ell.rs:54-67); new_with_matrix starts go to batch_warp.cu's kernel
            double m[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) m[j] = j == lane ? (a.diag0 ? a.diag0[pid * a.diag0_stride + lane] : 1.0) : 0.0;
            m_store(tm_m, m);
            tm_wait_st();
        }
        double xc = a.xc0[pid * a.xc0_stride + lane];
        double kappa = a.kappa0[pid * a.kappa0_stride];
        double gamma = a.gamma0 ? a.gamma0[pid * a.gamma0_stride] : 0.0;
        double tsq = 0.0;
        Lowpass32<kTabTm, kAligned> om;
        om.init(a, pid);

        bool has_best = false;
        double best = 0.0;
        long long niter = 0;
        int exit_status = kExitMaxIt;

        for (; niter < a.max_iters; ++niter) {  // cutting_plane_optim, src/cutting_plane.rs:299-311
            c.sx()[lane] = xc;
            __syncwarp();
            double g = 0.0;
            Cut cut;
            const bool shrunk = om.assess_optim(c, gamma, g, cut);
            if (shrunk) { has_best = true; best = xc; }
            const int which = shrunk ? kCentral : kBias;

            // ---- Ell::update_core (src/ell.rs:97-137)
            c.sg()[lane] = g;
            __syncwarp();
            double m[32];
            m_load(tm_m, m);
            double gt = 0.0;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const double2 gv = *reinterpret_cast<const double2*>(c.sg() + j);
                gt += m[j] * gv.x;
                gt += m[j + 1] * gv.y;
            }
            c.sp()[lane] = g * gt;
            c.sgt()[lane] = gt;
            __syncwarp();
            const double omega = seq_sum32(c.sp());
            tsq = kappa * omega;
            Rsd r;
            const int status = ell_calc(cc, which, cut, tsq, r);
            if (status != kSuccess) { exit_status = status; break; }
            const double rho_over_omega = r.rho / omega;
            xc -= rho_over_omega * gt;
            const double ratio = r.sigma / omega;
            const double rg = ratio * gt;
            // element (lane, j): hi = max(lane, j) supplies ratio*grad_t, lo = min(lane, j) supplies grad_t.  ratio*grad_t[j] is
            // recomputed from the broadcast grad_t[j] (same operands, same rounding as in lane j) instead of being exchanged through
            // shared memory: 34 fewer wavefronts per iteration on the unit that bounds this kernel, for 32 more DMULs.
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const double2 gv = *reinterpret_cast<const double2*>(c.sgt() + j);
                double u0, u1;
                if (j <= lane) u0 = rg * gv.x; else u0 = (ratio * gv.x) * gt;
                if (j + 1 <= lane) u1 = rg * gv.y; else u1 = (ratio * gv.y) * gt;
                m[j] -= u0;
                m[j + 1] -= u1;
            }
            kappa *= r.delta;
            if (a.no_defer_trick) {
#pragma unroll
                for (int j = 0; j < 32; ++j) m[j] *= kappa;
                kappa = 1.0;
            }
            m_store(tm_m, m);
            tm_wait_st();
            if (tsq < a.tolerance) { exit_status = kExitTol; break; }
        }

        if (lane == 0) {
            a.num_iters[pid] = niter;
            a.exit_status[pid] = exit_status;
            if (a.has_xbest) a.has_xbest[pid] = has_best ? 1 : 0;
            if (a.gamma) a.gamma[pid] = gamma;
            if (a.tsq) a.tsq[pid] = tsq;
            if (a.kappa) a.kappa[pid] = kappa;
        }
        if (a.xbest && has_best) a.xbest[pid * kN + lane] = best;
        if (a.xc) a.xc[pid * kN + lane] = xc;
        __syncwarp();
    }

    tm_free_all(tm_base, warp);
}

#ifdef __CUDACC__   // host launchers: left out when tests/cpp/emu_lowpass32.cpp compiles the kernel above for the host
template <int kTabTm, bool kAligned, int kWarps>
int launch_variant(const BatchArgs& a, cudaStream_t s, int sm_count) {
    const size_t tab = ((size_t)a.aux[0] * kTabStride + 1) & ~(size_t)1;
    const size_t smem = (tab + (size_t)kWarps * kScratch) * sizeof(double);
    auto kern = lowpass32_tmem_kernel<kTabTm, kAligned, kWarps>;
    ELL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks = sm_count;  // persistent: one CTA per SM (it owns the SM's 512 TMEM columns)
    const long long need = (a.count + kWarps - 1) / kWarps;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, kWarps * 32, smem, s>>>(a);
    ELL_CUDA_TRY(cudaGetLastError());
    return ELLCUDA_OK;
}

#endif  // __CUDACC__

}  // namespace

#ifdef __CUDACC__
int launch_batch_lowpass32(const BatchArgs& a, cudaStream_t s, int sm_count, long long* launches) {
    if (a.n != kN || a.aux[0] != 15 * kN) { set_error("lowpass32 kernel needs n = 32 and a 480-row spectrum"); return ELLCUDA_E_ARG; }
    if (a.mq0) { set_error("lowpass32 kernel: diagonal starts only"); return ELLCUDA_E_ARG; }
    const int v = ELL_ENV_INT("ELLCUDA_LOWPASS_VARIANT", 0);   // performance experiments only (read once); results are identical
    int rc;
    // Measured on B200 (tools/probe_tune.py, 16384 problems): spectrum entirely in shared memory 503 M iter/s (default);
    // last 8 columns of every row from TMEM 490 (the two-block select costs more issue slots than the shared-memory
    // wavefronts it saves); the same with block-aligned rounds (one TMEM load, no select, 5.04 rounds per call) 482;
    // 20 warps x 96 registers 395 (spills).  batch_warp.cu's kernel (M in registers): 445.
    if (v == 1) rc = launch_variant<8, false, 16>(a, s, sm_count);
    else if (v == 2) rc = launch_variant<8, true, 16>(a, s, sm_count);
    else rc = launch_variant<0, false, 16>(a, s, sm_count);
    if (rc == ELLCUDA_OK && launches) *launches += 1;
    return rc;
}

#endif  // __CUDACC__

}  // namespace ellcuda
