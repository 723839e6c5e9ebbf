// batch_ell64.cu — batched cutting-plane loops on `Ell` (src/ell.rs) for n <= 64: ONE WARP PER PROBLEM, persistent, with
// the shape matrix M in TENSOR MEMORY.
//
// SURVEY.md §2 K1 asks for "one warp (n <= 32) / warp-group (n = 64)"; batch_warp.cu covers n <= 32 with M's rows in
// registers (one row per lane).  For 32 < n <= 64 a lane owns rows L and L+32 — 128 doubles, too many for registers —
// so M lives in TMEM: a problem takes 256 of the SM's 512 32-bit columns in its warp's lane quarter (row L in columns
// [0,128), row L+32 in [128,256)), i.e. 8 problems per SM, and moves through registers in 16-double chunks
// (tcgen05.ld/st 32x32b.x32) for M·g and for the rank-1 update.  Shared memory then only holds the vectors, which
// leaves room for the LMI oracle's F tables.
//
//   cutting_plane_optim / _feas          src/cutting_plane.rs:286-313, 205-227
//   Ell::update_core                     src/ell.rs:97-137 (grad_t, omega, xc, lower-triangle rank-1 + mirror, kappa)
//   oracles: quadratic bowl (tests/integration_test.rs:90-105), LowpassOracle (src/oracles/lowpass_oracle.rs:7-150, table
//            read transposed from L2: n > 32 does not fit in shared memory), LMIOracle / LMIOldOracle / LMI0Oracle
//            (lmi_device.cuh) — the reference runs its LMI oracles under `Ell` as well (tests/lmi_tests.rs:174-225).
// Every sum keeps the reference's order (row dots j ascending from 0.0, omega k ascending), -fmad=false: results are
// bit-identical to the Rust path.
#include "common.cuh"
#include "lmi_device.cuh"
#include "tmem.cuh"

namespace ellcuda {
namespace {

#ifndef ELLCUDA_ELL64_WARPS
#define ELLCUDA_ELL64_WARPS 8        // the host emulation under ThreadSanitizer builds with 6 (it tracks <= 256 running threads)
#endif
constexpr int kWarps = ELLCUDA_ELL64_WARPS;   // 8 problems per SM: 4 lane quarters x 2 column halves
constexpr int kColsPerProblem = 256; // 2 rows x 64 doubles x 2 columns
constexpr int kVec = 5 * kNMax;      // per-warp vectors: sx | sg | sgt | srg | sp

struct Vecs {
    double* sx;   // xc (oracle input)
    double* sg;   // g
    double* sgt;  // grad_t = M g
    double* srg;  // ratio * grad_t
    double* sp;   // g .* grad_t / oracle scratch
};

// ---------------------------------------------------------------------------------------------- Ell in tensor memory
struct Ell64 {
    uint32_t tm;   // TMEM address of this problem's column base in this warp's lane quarter
    int n, nch;    // nch = number of 16-double chunks per row

    // Ell::new_with_matrix / new / new_with_scalar (src/ell.rs:31-78)
    __device__ void init(const BatchArgs& a, long long pid, int lane) {
        const int i0 = lane, i1 = lane + 32;
        for (int c = 0; c < nch; ++c) {
            uint32_t r0[32], r1[32];
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const int j = 16 * c + t;
                double v0 = 0.0, v1 = 0.0;
                if (j < n) {
                    if (a.mq0) {
                        if (i0 < n) v0 = a.mq0[pid * a.mq0_stride + (long long)i0 * n + j];
                        if (i1 < n) v1 = a.mq0[pid * a.mq0_stride + (long long)i1 * n + j];
                    } else {
                        if (j == i0) v0 = a.diag0 ? a.diag0[pid * a.diag0_stride + i0] : 1.0;
                        if (j == i1) v1 = a.diag0 ? a.diag0[pid * a.diag0_stride + i1] : 1.0;
                    }
                }
                r0[2 * t] = (uint32_t)__double2loint(v0); r0[2 * t + 1] = (uint32_t)__double2hiint(v0);
                r1[2 * t] = (uint32_t)__double2loint(v1); r1[2 * t + 1] = (uint32_t)__double2hiint(v1);
            }
            tm_st32(tm + 32 * c, r0);
            tm_st32(tm + 128 + 32 * c, r1);
        }
        tm_wait_st();
    }

    // Ell::update_core (src/ell.rs:97-137).  g0/g1 = this lane's entries of the gradient (0 beyond n).  `mirror_src`: the
    // problem's new_with_matrix input while no cut has succeeded yet (iteration 0), else nullptr — the first successful
    // cut rewrites the upper triangle from the still original lower one (:117-128), see batch_warp.cu.
    __device__ int update(const CalcConsts& cc, int which, const Cut& cut, const Vecs& v, int lane, double g0, double g1,
                          double& xc0, double& xc1, double& kappa, double& tsq, bool no_defer, const double* mirror_src) {
        const int i0 = lane, i1 = lane + 32;
        const bool on0 = i0 < n, on1 = i1 < n;
        v.sg[i0] = g0;
        v.sg[i1] = g1;
        __syncwarp();
        double gt0 = 0.0, gt1 = 0.0;   // grad_t[i] = sum_j M[i][j]*g[j], j ascending from 0.0 (src/arr.rs:426-442)
        for (int c = 0; c < nch; ++c) {
            uint32_t r0[32], r1[32];
            tm_ld32(tm + 32 * c, r0);
            tm_ld32(tm + 128 + 32 * c, r1);
            tm_wait_ld();
#pragma unroll
            for (int t = 0; t < 16; t += 2) {
                const int j = 16 * c + t;
                const double2 gv = *reinterpret_cast<const double2*>(v.sg + j);
                if (j < n) {
                    gt0 += mk_double(r0[2 * t], r0[2 * t + 1]) * gv.x;
                    gt1 += mk_double(r1[2 * t], r1[2 * t + 1]) * gv.x;
                }
                if (j + 1 < n) {
                    gt0 += mk_double(r0[2 * t + 2], r0[2 * t + 3]) * gv.y;
                    gt1 += mk_double(r1[2 * t + 2], r1[2 * t + 3]) * gv.y;
                }
            }
        }
        if (!on0) gt0 = 0.0;
        if (!on1) gt1 = 0.0;
        v.sp[i0] = g0 * gt0;
        v.sp[i1] = g1 * gt1;
        v.sgt[i0] = gt0;
        v.sgt[i1] = gt1;
        __syncwarp();
        double omega = 0.0;            // grad.dot(grad_t), k ascending (src/arr.rs:443-451); every lane, same order
        for (int k = 0; k < n; ++k) omega += v.sp[k];
        tsq = kappa * omega;
        Rsd r;
        const int status = ell_calc(cc, which, cut, tsq, r);
        if (status != kSuccess) return status;
        const double rho_over_omega = r.rho / omega;
        xc0 -= rho_over_omega * gt0;
        xc1 -= rho_over_omega * gt1;
        const double ratio = r.sigma / omega;
        const double rg0 = ratio * gt0, rg1 = ratio * gt1;
        v.srg[i0] = rg0;
        v.srg[i1] = rg1;
        __syncwarp();
        kappa *= r.delta;
        const double scale = kappa;
        if (no_defer) kappa = 1.0;     // mq *= kappa; kappa = 1 (:132-135)
        // element (i, j): hi = max(i, j) supplies ratio*grad_t, lo = min(i, j) supplies grad_t
        for (int c = 0; c < nch; ++c) {
            uint32_t r0[32], r1[32];
            tm_ld32(tm + 32 * c, r0);
            tm_ld32(tm + 128 + 32 * c, r1);
            tm_wait_ld();
#pragma unroll
            for (int t = 0; t < 16; t += 2) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j = 16 * c + t + h;
                    if (j < n) {
                        const double gtj = v.sgt[j], rgj = v.srg[j];
                        double m0 = mk_double(r0[2 * (t + h)], r0[2 * (t + h) + 1]);
                        double m1 = mk_double(r1[2 * (t + h)], r1[2 * (t + h) + 1]);
                        if (mirror_src) {
                            if (on0 && j > i0) m0 = mirror_src[(long long)j * n + i0];
                            if (on1 && j > i1) m1 = mirror_src[(long long)j * n + i1];
                        }
                        double u0, u1;
                        if (j <= i0) u0 = rg0 * gtj; else u0 = rgj * gt0;
                        if (j <= i1) u1 = rg1 * gtj; else u1 = rgj * gt1;
                        m0 -= u0;
                        m1 -= u1;
                        if (no_defer) { m0 *= scale; m1 *= scale; }
                        r0[2 * (t + h)] = (uint32_t)__double2loint(m0); r0[2 * (t + h) + 1] = (uint32_t)__double2hiint(m0);
                        r1[2 * (t + h)] = (uint32_t)__double2loint(m1); r1[2 * (t + h) + 1] = (uint32_t)__double2hiint(m1);
                    }
                }
            }
            tm_st32(tm + 32 * c, r0);
            tm_st32(tm + 128 + 32 * c, r1);
        }
        tm_wait_st();
        return status;
    }
};

// ---------------------------------------------------------------------------------------------- oracles
enum : int { kCut = 0, kShrunk = 1, kFeasible = 2 };

struct Common {
    const BatchArgs* a;
    int lane, n;
    Vecs v;
    // LMI only
    LmiShared lmi;
    WarpMem wm;
};

struct OrBowl64 {  // tests/integration_test.rs:90-105
    static constexpr bool kFeas = false;
    double t0, t1;
    __device__ void init(const Common& c, long long pid) {
        const double* p = c.a->params + pid * c.a->params_stride;
        t0 = c.lane < c.n ? p[c.lane] : 0.0;
        t1 = c.lane + 32 < c.n ? p[c.lane + 32] : 0.0;
    }
    __device__ int assess(const Common& c, double xc0, double xc1, double& gamma, double& g0, double& g1, Cut& cut) {
        const double d0 = xc0 - t0, d1 = xc1 - t1;
        g0 = 2.0 * d0;
        g1 = 2.0 * d1;
        c.v.sp[c.lane] = d0 * d0;
        c.v.sp[c.lane + 32] = d1 * d1;
        __syncwarp();
        double f = 0.0;
        for (int k = 0; k < c.n; ++k) f += c.v.sp[k];
        __syncwarp();
        cut = single_cut(f);
        if (f < gamma) { gamma = f; return kShrunk; }
        return kCut;
    }
};

struct OrLowpass64 {  // src/oracles/lowpass_oracle.rs:7-150; spectrum read transposed (tt[j*mdim + row]) through L1/L2
    static constexpr bool kFeas = false;
    int idx1, idx2, idx3, nwpass, nwstop, mdim, kmax;
    double lp_sq, up_sq, sp_sq, fmax;
    const double* tt;

    __device__ void init(const Common& c, long long pid) {  // :23-52
        const double* p = c.a->params + pid * c.a->params_stride;
        mdim = c.a->aux[0];
        tt = c.a->table_t;
        nwpass = (int)floor(p[0] * (double)(mdim - 1)) + 1;
        nwstop = (int)floor(p[1] * (double)(mdim - 1)) + 1;
        lp_sq = p[2]; up_sq = p[3]; sp_sq = p[4];
        idx1 = -1; idx2 = nwpass - 1; idx3 = nwstop - 1;
        fmax = -INFINITY; kmax = -1;
    }
    __device__ __forceinline__ void row_elems(const Common& c, int k, double& g0, double& g1, double sign) const {
        g0 = c.lane < c.n ? sign * __ldg(tt + (long long)c.lane * mdim + k) : 0.0;
        g1 = c.lane + 32 < c.n ? sign * __ldg(tt + (long long)(c.lane + 32) * mdim + k) : 0.0;
    }
    // one round-robin band sweep in rounds of 32 rows (lane L takes traversal position base + L), stopping at the first row in
    // traversal order that violates.  BAND: 0 pass-band (:62-78), 1 stop-band (:83-104), 2 transition band (:106-119).
    template <int BAND>
    __device__ __forceinline__ bool sweep(const Common& c, int lo, int hi, int& cursor, double& g0, double& g1, Cut& cut) {
        const int len = hi - lo;
        double best_v = -INFINITY;
        int best_c = 0x7fffffff, best_k = -1;
        for (int base = 0; base < len; base += 32) {
            const int cidx = base + c.lane;
            int kk = cursor + 1 + cidx;
            if (kk >= hi) kk -= len;
            const bool active = cidx < len;
            const int k = active ? kk : lo;           // inactive lanes read a valid row and are masked below
            double val = 0.0;                         // spectrum[k].dot(x), j ascending (src/arr.rs:443-451)
            const double* col = tt + k;
#pragma unroll 4
            for (int j = 0; j < c.n; ++j) val += __ldg(col + (long long)j * mdim) * c.v.sx[j];
            bool viol;
            if (BAND == 0) viol = active && (val > up_sq || val < lp_sq);
            else if (BAND == 1) viol = active && (val > sp_sq || val < 0.0);
            else viol = active && (val < 0.0);
            const unsigned mask = __ballot_sync(0xffffffffu, viol);
            if (mask) {
                const int first = __ffs(mask) - 1;
                const int kf = __shfl_sync(0xffffffffu, k, first);
                const double vf = __shfl_sync(0xffffffffu, val, first);
                cursor = kf;
                if (BAND == 0) {
                    if (vf > up_sq) { row_elems(c, kf, g0, g1, 1.0); cut = parallel_cut(vf - up_sq, vf - lp_sq); }
                    else { row_elems(c, kf, g0, g1, -1.0); cut = parallel_cut(-vf + lp_sq, -vf + up_sq); }
                } else if (BAND == 1) {
                    if (vf > sp_sq) { row_elems(c, kf, g0, g1, 1.0); cut = parallel_cut(vf - sp_sq, vf); }
                    else { row_elems(c, kf, g0, g1, -1.0); cut = parallel_cut(-vf, -vf + sp_sq); }
                } else {
                    row_elems(c, kf, g0, g1, -1.0);
                    cut = single_cut(-vf);            // ParallelCut(-val, None)
                }
                return true;
            }
            if (BAND == 1 && active && val > best_v) { best_v = val; best_c = cidx; best_k = k; }   // first occurrence of the maximum
        }
        int last = cursor + len;                       // no violation: the cursor sits on the last row visited
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
    __device__ int assess(const Common& c, double xc0, double xc1, double& gamma, double& g0, double& g1, Cut& cut) {  // :137-150
        (void)xc0; (void)xc1;
        sp_sq = gamma;
        if (sweep<0>(c, 0, nwpass, idx1, g0, g1, cut)) return kCut;
        fmax = -INFINITY; kmax = -1;
        if (sweep<1>(c, nwstop, mdim, idx3, g0, g1, cut)) return kCut;
        if (sweep<2>(c, nwpass, nwstop, idx2, g0, g1, cut)) return kCut;
        const double x0 = c.v.sx[0];
        if (x0 < 0.0) {  // :123-127
            g0 = c.lane == 0 ? -1.0 : 0.0;
            g1 = 0.0;
            cut = single_cut(-x0);
            return kCut;
        }
        row_elems(c, kmax, g0, g1, 1.0);
        cut = parallel_cut(0.0, fmax);
        gamma = fmax;
        return kShrunk;
    }
};

struct OrLmi64 {  // LMIOracle / LMIOldOracle / LMI0Oracle under cutting_plane_feas (lmi_device.cuh)
    static constexpr bool kFeas = true;
    __device__ void init(const Common&, long long) {}
    __device__ int assess(const Common& c, double, double, double&, double& g0, double& g1, Cut& cut) {
        double ep = 0.0;
        if (!lmi_assess(c.lmi, c.wm, c.lane, g0, g1, ep)) return kFeasible;
        cut = single_cut(ep);
        return kCut;
    }
};

// ---------------------------------------------------------------------------------------------- kernel
template <class Oracle>
__global__ void __launch_bounds__(kWarps * 32, 1) batch_ell64_kernel(const BatchArgs a, int fstride, int per_warp_doubles, int lmi0) {
    extern __shared__ __align__(16) double smem[];
    __shared__ uint32_t tmem_base_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = a.n;
    Common c;
    c.a = &a; c.lane = lane; c.n = n;
    long long base = 0;
    if (Oracle::kFeas) {
        const int m = a.aux[0], tri = m * (m + 1) / 2;
        base = lmi_stage_tables(a, smem, fstride, lmi0);
        c.lmi = LmiShared{smem, smem + (long long)n * fstride, fstride, m, tri, n, lmi0};
    }
    double* ws = smem + base + (long long)warp * per_warp_doubles;
    c.v = Vecs{ws, ws + kNMax, ws + 2 * kNMax, ws + 3 * kNMax, ws + 4 * kNMax};
    if (Oracle::kFeas) {
        const int m = a.aux[0], tri = m * (m + 1) / 2;
        c.wm.sx = c.v.sx; c.wm.sv = nullptr; c.wm.sw = nullptr; c.wm.sz = nullptr; c.wm.qu = nullptr;
        c.wm.sa = ws + kVec;
        c.wm.st = c.wm.sa + ((tri + 1) & ~1);
        c.wm.swit = c.wm.st + m * m;
    }
    const uint32_t tm_base = tm_alloc_all(&tmem_base_s, warp);
    Ell64 ell;
    ell.tm = tm_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * kColsPerProblem);
    ell.n = n;
    ell.nch = (n + 15) / 16;
    const CalcConsts cc = CalcConsts::make(n, a.use_parallel_cut);
    const int i0 = lane, i1 = lane + 32;
    const bool on0 = i0 < n, on1 = i1 < n;
    for (int k = lane; k < kVec; k += 32) ws[k] = 0.0;    // entries beyond n stay 0 (read as pairs by the M.g loop)
    __syncwarp();

    for (;;) {
        unsigned long long pid_u = 0;
        if (lane == 0) pid_u = atomicAdd(a.queue, 1ULL);
        pid_u = __shfl_sync(0xffffffffu, pid_u, 0);
        const long long pid = (long long)pid_u;
        if (pid >= a.count) break;

        ell.init(a, pid, lane);
        double xc0 = on0 ? a.xc0[pid * a.xc0_stride + i0] : 0.0;
        double xc1 = on1 ? a.xc0[pid * a.xc0_stride + i1] : 0.0;
        double kappa = a.kappa0[pid * a.kappa0_stride];
        double gamma = (!Oracle::kFeas && a.gamma0) ? a.gamma0[pid * a.gamma0_stride] : 0.0;
        double tsq = 0.0;
        Oracle om;
        om.init(c, pid);
        bool has_best = false;
        double best0 = 0.0, best1 = 0.0;
        long long niter = 0;
        int exit_status = kExitMaxIt;

        for (; niter < a.max_iters; ++niter) {   // cutting_plane_optim :299-311 / cutting_plane_feas :215-225
            c.v.sx[i0] = xc0;
            c.v.sx[i1] = xc1;
            __syncwarp();
            double g0 = 0.0, g1 = 0.0;
            Cut cut = single_cut(0.0);
            const int what = om.assess(c, xc0, xc1, gamma, g0, g1, cut);
            if (what == kFeasible) { has_best = true; best0 = xc0; best1 = xc1; exit_status = kExitFeasible; break; }
            if (what == kShrunk) { has_best = true; best0 = xc0; best1 = xc1; }
            if (!on0) g0 = 0.0;
            if (!on1) g1 = 0.0;
            const double* mirror = (a.mq0 != nullptr && niter == 0) ? a.mq0 + pid * a.mq0_stride : nullptr;
            const int status = ell.update(cc, what == kShrunk ? kCentral : kBias, cut, c.v, lane, g0, g1, xc0, xc1, kappa, tsq,
                                          a.no_defer_trick != 0, mirror);
            if (status != kSuccess) { exit_status = status; break; }
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
        if (on0) {
            if (a.xbest && has_best) a.xbest[pid * n + i0] = best0;
            if (a.xc) a.xc[pid * n + i0] = xc0;
        }
        if (on1) {
            if (a.xbest && has_best) a.xbest[pid * n + i1] = best1;
            if (a.xc) a.xc[pid * n + i1] = xc1;
        }
        __syncwarp();
    }
    tm_free_all(tm_base, warp);
}

#ifdef __CUDACC__   // host launchers (and the transpose they launch): left out when tests/cpp/emu_ell64.cpp compiles the kernel above for the host
__global__ void transpose_table_kernel(const double* __restrict__ in, double* __restrict__ out, int rows, int cols) {
    __shared__ double t[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = by + r, j = bx + threadIdx.x;
        if (i < rows && j < cols) t[r][threadIdx.x] = in[(long long)i * cols + j];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = bx + r, i = by + threadIdx.x;   // out[j][i] = in[i][j]
        if (i < rows && j < cols) out[(long long)j * rows + i] = t[threadIdx.x][r];
    }
}

template <class Oracle>
int launch_k(const BatchArgs& a, cudaStream_t s, int sm_count, size_t fixed_bytes, int fstride, int per_warp, int lmi0) {
    const size_t smem = fixed_bytes + (size_t)kWarps * per_warp * 8;
    if (smem > 227 * 1024) { set_error("oracle tables do not fit in shared memory"); return ELLCUDA_E_ARG; }
    auto kern = batch_ell64_kernel<Oracle>;
    ELL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks = sm_count;   // persistent: one CTA per SM (it owns the SM's tensor memory)
    const long long need = (a.count + kWarps - 1) / kWarps;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, kWarps * 32, smem, s>>>(a, fstride, per_warp, lmi0);
    ELL_CUDA_TRY(cudaGetLastError());
    return ELLCUDA_OK;
}

#endif  // __CUDACC__

}  // namespace

#ifdef __CUDACC__
// out[j][i] = in[i][j] for a rows x cols table (the lowpass spectrum for the kernels that read it from L2)
int launch_transpose_table(const double* in, double* out, int rows, int cols, cudaStream_t s) {
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32)), block(32, 8);
    transpose_table_kernel<<<grid, block, 0, s>>>(in, out, rows, cols);
    ELL_CUDA_TRY(cudaGetLastError());
    return ELLCUDA_OK;
}

// Ell (KIND_ELL) with n <= 64: bowl / lowpass under cutting_plane_optim, the LMI oracles under cutting_plane_feas
int launch_batch_ell64(const BatchArgs& a_in, int oracle_id, cudaStream_t s, int sm_count, long long* launches) {
    BatchArgs a = a_in;
    if (a.kind != ELL_KIND_ELL) { set_error("this kernel runs the Ell search space"); return ELLCUDA_E_ARG; }
    if (a.n < 1 || a.n > kNMax) { set_error("batched Ell needs 1 <= n <= 64"); return ELLCUDA_E_ARG; }
    int rc;
    if (oracle_id == ELL_ORACLE_BOWL) {
        if (a.mode != ELL_MODE_OPTIM) { set_error("BOWL runs under cutting_plane_optim"); return ELLCUDA_E_ARG; }
        if (!a.params || a.params_width < a.n) { set_error("BOWL needs params[n] = target"); return ELLCUDA_E_ARG; }
        rc = launch_k<OrBowl64>(a, s, sm_count, 0, 0, kVec, 0);
    } else if (oracle_id == ELL_ORACLE_LOWPASS) {
        if (a.mode != ELL_MODE_OPTIM) { set_error("LOWPASS runs under cutting_plane_optim"); return ELLCUDA_E_ARG; }
        if (!a.params || a.params_width < 5 || !a.table || a.aux[0] != 15 * a.n || a.table_len < (long long)a.aux[0] * a.n) {
            set_error("LOWPASS needs params[5], the spectrum table and aux[0] = 15*n");
            return ELLCUDA_E_ARG;
        }
        double* tt = nullptr;   // the spectrum transposed: lanes evaluating consecutive rows then read consecutive addresses
        ELL_CUDA_TRY(cudaMallocAsync((void**)&tt, (size_t)a.aux[0] * a.n * 8, s));
        rc = launch_transpose_table(a.table, tt, a.aux[0], a.n, s);
        if (rc == ELLCUDA_OK) {
            a.table_t = tt;
            if (launches) *launches += 1;
            rc = launch_k<OrLowpass64>(a, s, sm_count, 0, 0, kVec, 0);
        }
        cudaFreeAsync(tt, s);
    } else if (oracle_id == ELL_ORACLE_LMI || oracle_id == ELL_ORACLE_LMI_OLD || oracle_id == ELL_ORACLE_LMI0) {
        const int lmi0 = oracle_id == ELL_ORACLE_LMI0;
        if (a.mode != ELL_MODE_FEAS) { set_error("the LMI device oracles run under cutting_plane_feas"); return ELLCUDA_E_ARG; }
        const int n = a.n, m = a.aux[0];
        if (m < 1 || m > 32) { set_error("LMI kernel needs 1 <= m <= 32"); return ELLCUDA_E_ARG; }
        if (!a.table || a.table_len < (long long)(n + (lmi0 ? 0 : 1)) * m * m) {
            set_error(lmi0 ? "LMI0 needs table = F_0..F_{n-1}" : "LMI needs table = F_0..F_{n-1}, B");
            return ELLCUDA_E_ARG;
        }
        const int tri = m * (m + 1) / 2, fstride = tri | 1;
        int per_warp = kVec + ((tri + 1) & ~1) + m * m + ((m + 1) & ~1);
        per_warp = (per_warp + 1) & ~1;
        const size_t fixed = (((size_t)n * fstride + tri + 1) & ~(size_t)1) * 8;
        rc = launch_k<OrLmi64>(a, s, sm_count, fixed, fstride, per_warp, lmi0);
    } else {
        set_error("oracle not available on the batched Ell (n <= 64) kernel");
        return ELLCUDA_E_ARG;
    }
    if (rc == ELLCUDA_OK && launches) *launches += 1;
    return rc;
}

#endif  // __CUDACC__

}  // namespace ellcuda
