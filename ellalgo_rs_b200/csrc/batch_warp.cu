// batch_warp.cu — batched cutting-plane loop for 3 <= n <= 32: ONE WARP PER PROBLEM, persistent.
//
// Layout: lane i owns row i of the shape matrix M in registers (32 doubles) and component i of xc / xbest; the
// per-iteration vectors (xc, g, grad_t, ratio*grad_t, g.*grad_t) are exchanged through a 1.25 KB per-warp
// shared-memory scratch read back as broadcasts.  Warps pull problems from a global counter, so iteration
// counts that differ by orders of magnitude between problems do not idle the SM.
//
// Arithmetic is the reference's, in the reference's order (compiled -fmad=false):
//   grad_t[i] = sum_j M[i][j]*g[j]   left to right from 0.0          src/arr.rs:426-442  (lane i, sequential)
//   omega     = sum_k g[k]*grad_t[k] left to right                   src/arr.rs:443-451  (every lane, same order)
//   M[i][j]  -= (ratio*grad_t[max(i,j)]) * grad_t[min(i,j)]          src/ell.rs:117-128  (lower triangle + mirror)
// so num_iters and the exit status are identical to the Rust path, not merely close.
// Drivers: cutting_plane_optim (src/cutting_plane.rs:286-313).
// Oracles: shifted quadratic bowl (tests/integration_test.rs:90-105), LowpassOracle (src/oracles/lowpass_oracle.rs).
#include <cstdlib>

#include "common.cuh"

namespace ellcuda {
namespace {

constexpr int kWarpScratch = 5 * 32;  // doubles per warp: sx, sg, sgt, srg, sp

struct WarpCtx {
    int lane;
    int n;
    double* ws;         // per-warp scratch: sx | sg | sgt | srg | sp, 32 doubles each
    const double* tab;  // shared-memory copy of the oracle table (padded rows) or nullptr
    int tab_stride;
    __device__ __forceinline__ double* sx() const { return ws; }        // xc
    __device__ __forceinline__ double* sg() const { return ws + 32; }   // g
    __device__ __forceinline__ double* sgt() const { return ws + 64; }  // grad_t
    __device__ __forceinline__ double* srg() const { return ws + 96; }  // ratio * grad_t
    __device__ __forceinline__ double* sp() const { return ws + 128; }  // g .* grad_t (oracle scratch before the update)
};

template <bool FIXED32>
__device__ __forceinline__ double seq_sum(const double* s, int n) {
    double acc = 0.0;
    if constexpr (FIXED32) {
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
            const double2 v = *reinterpret_cast<const double2*>(s + k);
            acc += v.x;
            acc += v.y;
        }
    } else {
        for (int k = 0; k < n; ++k) acc += s[k];
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------- oracles
template <bool FIXED32>
struct OrBowl {  // tests/integration_test.rs:90-105
    double t;
    __device__ void init(const BatchArgs& a, long long pid, const WarpCtx& c) {
        t = (c.lane < c.n) ? a.params[pid * a.params_stride + c.lane] : 0.0;
    }
    __device__ bool assess_optim(const WarpCtx& c, double xc, double& gamma, double& g, Cut& cut) {
        const double diff = xc - t;
        g = 2.0 * diff;
        c.sp()[c.lane] = diff * diff;
        __syncwarp();
        const double f = seq_sum<FIXED32>(c.sp(), c.n);
        cut = single_cut(f);
        if (f < gamma) {
            gamma = f;
            return true;
        }
        return false;
    }
    static constexpr bool kNeedsTable = false;
};

template <bool FIXED32, int kRowsPerLane, bool XREG = false, int kUnroll = 4>
struct OrLowpass {  // src/oracles/lowpass_oracle.rs:7-150
    int idx1, idx2, idx3, nwpass, nwstop, mdim, kmax;
    double lp_sq, up_sq, sp_sq, fmax;
    static constexpr bool kNeedsTable = true;

    __device__ void init(const BatchArgs& a, long long pid, const WarpCtx& c) {  // :23-52
        const double* p = a.params + pid * a.params_stride;
        mdim = a.aux[0];
        nwpass = (int)floor(p[0] * (double)(mdim - 1)) + 1;
        nwstop = (int)floor(p[1] * (double)(mdim - 1)) + 1;
        lp_sq = p[2]; up_sq = p[3]; sp_sq = p[4];
        idx1 = -1; idx2 = nwpass - 1; idx3 = nwstop - 1;
        fmax = -INFINITY; kmax = -1;
        (void)c;
    }
    // spectrum[k].dot(x) for the R rows this lane evaluates in one round: each sum is sequential over j
    // (src/arr.rs:443-451); x is the broadcast copy in c.sx(); the R chains are independent and interleave.
    template <int R>
    __device__ __forceinline__ void row_dots(const WarpCtx& c, const int (&k)[R], double (&acc)[R], const double (&xr)[32]) const {
        const double* row[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            row[r] = c.tab + (long long)k[r] * c.tab_stride;
            acc[r] = 0.0;
        }
        if constexpr (FIXED32 && XREG) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] += row[r][j] * xr[j];
            }
        } else if constexpr (FIXED32) {
#pragma unroll kUnroll
            for (int j = 0; j < 32; j += 2) {
                const double2 x = *reinterpret_cast<const double2*>(c.sx() + j);
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] += row[r][j] * x.x;
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] += row[r][j + 1] * x.y;
            }
        } else {
            for (int j = 0; j < c.n; ++j) {
                const double x = c.sx()[j];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] += row[r][j] * x;
            }
        }
    }
    __device__ __forceinline__ double row_elem(const WarpCtx& c, int k) const {
        return c.lane < c.n ? c.tab[(long long)k * c.tab_stride + c.lane] : 0.0;
    }

    // One round-robin band sweep, 32·R rows per round (R per lane; traversal position = base + 32·r + lane), stopping
    // at the first row in traversal order that violates.  BAND: 0 pass-band (:62-78), 1 stop-band (:83-104),
    // 2 transition band (:106-119).  Returns true when a cut was produced.
    template <int BAND, int R>
    __device__ __forceinline__ bool sweep(const WarpCtx& c, int lo, int hi, int& cursor, double& g, Cut& cut, const double (&xr)[32]) {
        const int len = hi - lo;
        double best_v = -INFINITY;
        int best_c = 0x7fffffff, best_k = -1;
        for (int base = 0; base < len; base += 32 * R) {
            int k[R];
            double val[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int cidx = base + 32 * r + c.lane;
                int kk = cursor + 1 + cidx;
                if (kk >= hi) kk -= len;
                k[r] = cidx < len ? kk : lo;      // inactive slots read a valid row and are masked below
            }
            row_dots<R>(c, k, val, xr);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int cidx = base + 32 * r + c.lane;
                const bool active = cidx < len;
                bool viol;
                if (BAND == 0) viol = active && (val[r] > up_sq || val[r] < lp_sq);
                else if (BAND == 1) viol = active && (val[r] > sp_sq || val[r] < 0.0);
                else viol = active && (val[r] < 0.0);
                const unsigned mask = __ballot_sync(0xffffffffu, viol);
                if (mask) {
                    const int first = __ffs(mask) - 1;
                    const int kf = __shfl_sync(0xffffffffu, k[r], first);
                    const double vf = __shfl_sync(0xffffffffu, val[r], first);
                    cursor = kf;
                    const double e = row_elem(c, kf);
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
                    if (active && val[r] > best_v) { best_v = val[r]; best_c = cidx; best_k = k[r]; }
                }
            }
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

    __device__ bool assess_optim(const WarpCtx& c, double xc, double& gamma, double& g, Cut& cut) {  // :137-150
        (void)xc;
        sp_sq = gamma;
        double xr[32];
        if constexpr (FIXED32 && XREG) {   // x once per oracle call instead of once per round
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const double2 v = *reinterpret_cast<const double2*>(c.sx() + j);
                xr[j] = v.x; xr[j + 1] = v.y;
            }
        }
        if (sweep<0, kRowsPerLane>(c, 0, nwpass, idx1, g, cut, xr)) return false;
        fmax = -INFINITY; kmax = -1;
        if (sweep<1, kRowsPerLane>(c, nwstop, mdim, idx3, g, cut, xr)) return false;
        if (sweep<2, kRowsPerLane>(c, nwpass, nwstop, idx2, g, cut, xr)) return false;
        const double x0 = c.sx()[0];
        if (x0 < 0.0) {  // :123-127
            g = c.lane == 0 ? -1.0 : 0.0;
            cut = single_cut(-x0);
            return false;
        }
        g = row_elem(c, kmax);
        cut = parallel_cut(0.0, fmax);
        gamma = fmax;
        return true;
    }
};

// ---------------------------------------------------------------------------------------------- kernel
template <class Oracle, bool FIXED32, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) batch_warp_kernel(const BatchArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n = FIXED32 ? 32 : a.n;

    WarpCtx c;
    c.lane = lane;
    c.n = n;
    double* tab_s = nullptr;
    int tab_stride = 0;
    double* scratch = smem;
    if (Oracle::kNeedsTable) {  // stage the shared table with odd row stride: lane L reading row k+L is conflict-free
        const int rows = a.aux[0];
        tab_stride = (n & 1) ? n : n + 1;
        tab_s = smem;
        for (int idx = threadIdx.x; idx < rows * n; idx += blockDim.x) {
            const int r = idx / n, col = idx - r * n;
            tab_s[r * tab_stride + col] = a.table[idx];
        }
        long long off = (long long)rows * tab_stride;
        off = (off + 1) & ~1LL;  // keep the scratch 16-byte aligned
        scratch = smem + off;
        __syncthreads();
    }
    c.tab = tab_s;
    c.tab_stride = tab_stride;
    c.ws = scratch + warp * kWarpScratch;

    const CalcConsts cc = CalcConsts::make(n, a.use_parallel_cut);
    const bool lane_on = lane < n;

    for (;;) {
        unsigned long long pid_u = 0;
        if (lane == 0) pid_u = atomicAdd(a.queue, 1ULL);
        pid_u = __shfl_sync(0xffffffffu, pid_u, 0);
        const long long pid = (long long)pid_u;
        if (pid >= a.count) break;

        // ---- Ell::new_with_matrix / new / new_with_scalar (src/ell.rs:31-78)
        double m[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            double v = 0.0;
            if (lane_on && j < n) {
                if (a.mq0) v = a.mq0[pid * a.mq0_stride + (long long)lane * n + j];
                else if (j == lane) v = a.diag0 ? a.diag0[pid * a.diag0_stride + lane] : 1.0;
            }
            m[j] = v;
        }
        double xc = lane_on ? a.xc0[pid * a.xc0_stride + lane] : 0.0;
        double kappa = a.kappa0[pid * a.kappa0_stride];
        double gamma = a.gamma0 ? a.gamma0[pid * a.gamma0_stride] : 0.0;
        double tsq = 0.0;
        Oracle om;
        om.init(a, pid, c);

        bool has_best = false;
        double best = 0.0;
        long long niter = 0;
        int exit_status = kExitMaxIt;
        // Ell::new_with_matrix takes any matrix (src/ell.rs:31-41): M.g uses the rows as given (:102) until the first successful
        // cut, whose rank-1 loop rewrites the upper triangle from the lower one (:117-128).  Until then M is still the input,
        // so that mirror is a re-read of the input's lower triangle (a no-op in value for a symmetric input).  The driver
        // returns at the first cut that is not a Success (src/cutting_plane.rs:308), so "first successful cut" = iteration 0.

        for (; niter < a.max_iters; ++niter) {  // cutting_plane_optim, src/cutting_plane.rs:299-311
            c.sx()[lane] = xc;
            __syncwarp();
            double g = 0.0;
            Cut cut;
            const bool shrunk = om.assess_optim(c, xc, gamma, g, cut);
            if (!lane_on) g = 0.0;
            if (shrunk) { has_best = true; best = xc; }
            const int which = shrunk ? kCentral : kBias;

            // ---- Ell::update_core (src/ell.rs:97-137)
            c.sg()[lane] = g;
            __syncwarp();
            double gt = 0.0;
            if constexpr (FIXED32) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const double2 gv = *reinterpret_cast<const double2*>(c.sg() + j);
                    gt += m[j] * gv.x;
                    gt += m[j + 1] * gv.y;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < n) gt += m[j] * c.sg()[j];
            }
            c.sp()[lane] = g * gt;
            c.sgt()[lane] = gt;
            __syncwarp();
            const double omega = seq_sum<FIXED32>(c.sp(), n);
            tsq = kappa * omega;
            Rsd r;
            const int status = ell_calc(cc, which, cut, tsq, r);
            if (status != kSuccess) { exit_status = status; break; }
            if (a.mq0 != nullptr && niter == 0) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (lane_on && j > lane && j < n) m[j] = a.mq0[pid * a.mq0_stride + (long long)j * n + lane];
            }
            const double rho_over_omega = r.rho / omega;
            xc -= rho_over_omega * gt;
            const double ratio = r.sigma / omega;
            const double rg = ratio * gt;
            c.srg()[lane] = rg;
            __syncwarp();
            // element (lane, j): hi = max(lane, j) supplies ratio*grad_t, lo = min(lane, j) supplies grad_t
            if constexpr (FIXED32) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const double2 gv = *reinterpret_cast<const double2*>(c.sgt() + j);
                    const double2 rv = *reinterpret_cast<const double2*>(c.srg() + j);
                    double u0, u1;
                    if (j <= lane) u0 = rg * gv.x; else u0 = rv.x * gt;
                    if (j + 1 <= lane) u1 = rg * gv.y; else u1 = rv.y * gt;
                    m[j] -= u0;
                    m[j + 1] -= u1;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (j < n) {
                        double u;
                        if (j <= lane) u = rg * c.sgt()[j]; else u = c.srg()[j] * gt;
                        m[j] -= u;
                    }
                }
            }
            kappa *= r.delta;
            if (a.no_defer_trick) {
#pragma unroll
                for (int j = 0; j < 32; ++j) m[j] *= kappa;
                kappa = 1.0;
            }
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
        if (lane_on) {
            if (a.xbest && has_best) a.xbest[pid * n + lane] = best;
            if (a.xc) a.xc[pid * n + lane] = xc;
        }
        __syncwarp();
    }
}

#ifdef __CUDACC__   // host launchers: left out when tests/cpp/emu_batch_warp.cpp compiles the kernels above for the host
template <class Oracle, bool FIXED32, int WARPS, int MINB>
int launch_one(const BatchArgs& a, cudaStream_t s, int sm_count) {
    const int n = a.n;
    size_t smem = (size_t)WARPS * kWarpScratch * sizeof(double);
    if (Oracle::kNeedsTable) {
        const int stride = (n & 1) ? n : n + 1;
        size_t t = (size_t)a.aux[0] * stride;
        t = (t + 1) & ~(size_t)1;
        smem += t * sizeof(double);
    }
    if (smem > 227 * 1024) { set_error("oracle table does not fit in shared memory"); return ELLCUDA_E_ARG; }
    auto kern = batch_warp_kernel<Oracle, FIXED32, WARPS, MINB>;
    ELL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    ELL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    long long blocks = (long long)sm_count * per_sm;  // persistent: one wave exactly
    const long long need = (a.count + WARPS - 1) / WARPS;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, WARPS * 32, smem, s>>>(a);
    ELL_CUDA_TRY(cudaGetLastError());
    return ELLCUDA_OK;
}

// ELLCUDA_TUNE=<k> selects an alternative warps-per-CTA / registers-per-thread variant (performance experiments only;
// every variant computes identical results)
int tuning_variant() {
    return ELL_ENV_INT("ELLCUDA_TUNE", 0);   // read once per process
}

#endif  // __CUDACC__

}  // namespace

#ifdef __CUDACC__
int launch_batch_warp(const BatchArgs& a, int oracle_id, cudaStream_t s, int sm_count, long long* launches) {
    // Above n = 32 the CTA-per-problem kernel (packed lower triangle in shared memory: half the rank-1 work, more problems
    // resident) beats the warp-per-problem TMEM kernel — bowl n = 33 / 48 / 64: 389 / 285 / 218 against 217 / 190 / 143 M
    // iterations/s; lowpass n = 48 / 56 / 64: 71 / 37 / 30 against 50 / 29 / 15, but n = 36 / 44: 95 / 52 against 110 / 72.
    // ELLCUDA_CTA_MIN_N overrides the smallest n that goes to it (performance experiments; identical results).
    int cta_min = ELL_ENV_INT("ELLCUDA_CTA_MIN_N", 0);
    if (cta_min <= 0) cta_min = oracle_id == ELL_ORACLE_LOWPASS ? 48 : 33;
    if (a.n > 32 && a.n >= cta_min && a.n <= 224 && a.kind == ELL_KIND_ELL && a.mode == ELL_MODE_OPTIM)
        return launch_batch_ellcta(a, oracle_id, s, sm_count, launches);   // packed lower triangle of M in shared memory, one CTA
    if (a.n > 32 && a.n <= 64 && a.kind == ELL_KIND_ELL) return launch_batch_ell64(a, oracle_id, s, sm_count, launches);   // M in TMEM
    if (a.n < 1 || a.n > 224) { set_error("the batched Ell kernels need 1 <= n <= 224"); return ELLCUDA_E_ARG; }
    if (a.kind != ELL_KIND_ELL || a.mode != ELL_MODE_OPTIM) {
        set_error("warp-per-problem kernel: Ell + cutting_plane_optim only");
        return ELLCUDA_E_ARG;
    }
    int rc;
    if (oracle_id == ELL_ORACLE_BOWL) {
        if (!a.params || a.params_width < a.n) { set_error("BOWL needs params[n] = target"); return ELLCUDA_E_ARG; }
        const int v = tuning_variant();
        if (a.n != 32) rc = launch_one<OrBowl<false>, false, 8, 2>(a, s, sm_count);
        else if (v == 1) rc = launch_one<OrBowl<true>, true, 6, 2>(a, s, sm_count);   // 164 regs, no spills: 1038 M iter/s
        // default 8 warps x 2 CTAs (128 regs): 1100 M iter/s; 10x2 875; 12x2 and 16x1 no better
        else rc = launch_one<OrBowl<true>, true, 8, 2>(a, s, sm_count);
    } else if (oracle_id == ELL_ORACLE_LOWPASS) {
        if (!a.params || a.params_width < 5 || !a.table || a.aux[0] != 15 * a.n || a.table_len < (long long)a.aux[0] * a.n) {
            set_error("LOWPASS needs params[5], the spectrum table and aux[0] = 15*n");
            return ELLCUDA_E_ARG;
        }
        const int v = tuning_variant();
        // Measured on B200 (tools/probe_tune.py, 16384 problems; R = rows per lane per round, W = warps per CTA, U = unroll
        // of the row-dot loop):  R1 W16 U8  448 M iter/s (128 regs, 44 B spilled)  <- default
        //   R1 W16 U4 436 | U2 428 | U1 398 | U16 354 (970 B spilled) | R1 W12 U4 418 | R1 W8 x-in-registers 412 (196 regs)
        //   R1 W8 U16 404 | R2 W8 396 | R3 W8 363 | R4 W8 317 | 9-10 warps are allocated like 12 (168 regs, spills): 274-296.
        // The kernel is bound by shared-memory bandwidth (spectrum rows: 256 B per lane per row), so extra speculative
        // rows (R>1) cost more than the latency they hide; what pays is more resident warps without spilling.
        if (a.n != 32) rc = launch_one<OrLowpass<false, 1>, false, 8, 1>(a, s, sm_count);
        // TMEM-assisted kernel, batch_lowpass.cu (diagonal starts; a new_with_matrix start, which may be non-symmetric and then
        // needs the first-cut mirror below, runs in this file's kernel: 445 instead of 503 M iter/s)
        else if (v == 0 && !a.mq0) return launch_batch_lowpass32(a, s, sm_count, launches);
        else if (v == 1) rc = launch_one<OrLowpass<true, 2, false, 16>, true, 8, 1>(a, s, sm_count);
        else if (v == 2) rc = launch_one<OrLowpass<true, 1, true>, true, 8, 1>(a, s, sm_count);
        else rc = launch_one<OrLowpass<true, 1, false, 8>, true, 16, 1>(a, s, sm_count);
    } else {
        set_error("oracle not available in the warp-per-problem kernel");
        return ELLCUDA_E_ARG;
    }
    if (rc == ELLCUDA_OK && launches) *launches += 1;
    return rc;
}

#endif  // __CUDACC__

}  // namespace ellcuda
