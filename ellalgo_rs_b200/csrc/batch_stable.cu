// batch_stable.cu — batched cutting_plane_feas with EllStable and the LMI oracle: ONE WARP PER PROBLEM, persistent.
//
//   cutting_plane_feas                 src/cutting_plane.rs:205-227
//   EllStable::update_core             src/ell_stable.rs:52-125      (LDLᵀ-form update)
//   LMIOracle::assess_feas             src/oracles/lmi_oracle.rs:24-44
//   LDLTMgr::factor/witness/sym_quad   src/oracles/ldlt_mgr.rs:29-55, 98-111, 115-124
//
// n <= 64 variables (lane L owns indices L and L+32), LMI size m <= 32.  The F_k matrices (shared by the batch) are
// staged once per CTA in shared memory as packed lower triangles with an odd stride.  Every sum keeps the
// reference's order: the forward substitution is done column by column (for each j, every i>j subtracts term j —
// per element that is j ascending, as in the reference), the back substitution's per-row chain is sequential, the
// LDLᵀ inner sums accumulate k ascending per column.  Two variants of the space:
//   GENERAL  the full recurrence; strict upper triangle (the L factor) packed in shared memory (new_with_matrix)
//   DIAG     EllStable::new / new_with_scalar start from a diagonal matrix, and under the reference's recurrence the
//            strict upper triangle then stays exactly +0 forever: every O(n²) term is a multiply by
//            zero, so the same results come from the O(n) diagonal recurrence.  Bit-identical for finite data (the
//            LMI gradient sym_quad never yields -0.0; a non-finite gradient would differ in NaN placement only).
// The strict lower triangle of the reference's matrix is scratch that no API can observe (EllStable's fields are
// private, src/ell_stable.rs:9-15); it is recomputed on the fly instead of stored.
#include <cstdlib>

#include "common.cuh"
#include "lmi_device.cuh"
#include "stable_large_device.cuh"

namespace ellcuda {
namespace {

// ---------------------------------------------------------------------------------------------- EllStable update
// sequential sum of n values held as (lo, hi) per lane, via shared memory; every lane gets the result
__device__ __forceinline__ double seq_sum_pair(double* s, double lo, double hi, int lane, int n) {
    s[lane] = lo;
    if (lane + 32 < kNMax) s[lane + 32] = hi;
    __syncwarp();
    double acc = 0.0;
    int k = 0;
    for (; k + 1 < n; k += 2) {   // two per load, still one ordered chain
        const double2 v = *reinterpret_cast<const double2*>(s + k);
        acc += v.x;
        acc += v.y;
    }
    if (k < n) acc += s[k];
    __syncwarp();
    return acc;
}

template <bool DIAG>
__device__ int stable_update(const CalcConsts& cc, int which, const WarpMem& wm, int lane, int n, double g0, double g1,
                             const Cut& cut, double& d0, double& d1, double& xc0, double& xc1, double& kappa, double& tsq) {
    const int i0 = lane, i1 = lane + 32;
    const bool on0 = i0 < n, on1 = i1 < n;
    // --- inv(L)*grad (src/ell_stable.rs:60-69)
    double w0 = g0, w1 = g1;
    if (!DIAG) {
        for (int j = 0; j + 1 < n; ++j) {
            const double wj = bcast(w0, w1, j);
            const double* row = wm.qu + upper_off(j, n) - j;      // row[c] = Q[j][c], c >= j
            if (on0 && i0 > j) w0 -= row[i0] * wj;
            if (on1 && i1 > j) w1 -= row[i1] * wj;
        }
        d0 = on0 ? wm.qu[upper_off(i0, n)] : 0.0;
        d1 = on1 ? wm.qu[upper_off(i1, n)] : 0.0;
    }
    // --- inv(D)*inv(L)*grad and omega (:72-83)
    const double z0 = w0 * d0, z1 = w1 * d1;
    const double gg0 = z0 * w0, gg1 = z1 * w1;
    const double omega = seq_sum_pair(wm.sv, on0 ? gg0 : 0.0, on1 ? gg1 : 0.0, lane, n);   // sv now holds gg[]
    tsq = kappa * omega;
    Rsd r;
    const int status = ell_calc(cc, which, cut, tsq, r);
    if (status != kSuccess) return status;
    // --- mq*grad = inv(L')*inv(D)*inv(L)*grad (:93-98): q[c] = z[c] - sum_{j>c, ascending} (Q[c][j]*w[c])*q[j]
    double q0 = z0, q1 = z1;
    if (!DIAG) {
        wm.sw[lane] = w0;
        if (i1 < kNMax) wm.sw[i1] = w1;
        __syncwarp();
        for (int c = n - 2; c >= 0; --c) {
            const double wc = wm.sw[c];
            const double* row = wm.qu + upper_off(c, n) - c;
            if (on0 && i0 > c) wm.sz[i0] = (row[i0] * wc) * q0;   // the scratch entry val(j,c) = Q[c][j]*w[c], times q[j]
            if (on1 && i1 > c) wm.sz[i1] = (row[i1] * wc) * q1;
            __syncwarp();
            double qc = bcast(q0, q1, c);
            for (int j = c + 1; j < n; ++j) qc -= wm.sz[j];
            if (i0 == c) q0 = qc;
            if (i1 == c) q1 = qc;
            __syncwarp();
        }
    }
    // --- xc (:101-104)
    const double rho_over_omega = r.rho / omega;
    xc0 -= rho_over_omega * q0;
    xc1 -= rho_over_omega * q1;
    // --- rank-one update of the factor (:107-121)
    const double mu = r.sigma / (1.0 - r.sigma);
    double oldt = omega / mu;
    double oldt0 = 0.0, t0 = 1.0, oldt1 = 0.0, t1 = 1.0;
    {   // temp = oldt + gg[j]; oldt = temp — one ordered chain, run by every lane, the running values parked in sz so that each
        // lane picks its two afterwards (selecting inside the loop cost 5 instructions per step instead of 1)
        const double oldt_init = oldt;
        __syncwarp();
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const double temp = oldt + wm.sv[j];
            if (lane == 0) wm.sz[j] = temp;
            oldt = temp;
        }
        __syncwarp();
        if (on0) { oldt0 = i0 == 0 ? oldt_init : wm.sz[i0 - 1]; t0 = wm.sz[i0]; }
        if (on1) { oldt1 = wm.sz[i1 - 1]; t1 = wm.sz[i1]; }
        __syncwarp();
    }
    const double nd0 = d0 * (oldt0 / t0), nd1 = d1 * (oldt1 / t1);   // mq[j][j] *= oldt / temp
    if (DIAG) {
        d0 = nd0;
        d1 = nd1;
    } else {
        const double b0 = z0 / t0, b1 = z1 / t1;                       // beta2 = inv_md_inv_ml_g[j] / temp
        __syncwarp();
        wm.sz[lane] = b0;
        if (i1 < kNMax) wm.sz[i1] = b1;
        __syncwarp();
        for (int j = 0; j + 1 < n; ++j) {                              // mq[j][l] += beta2 * mq[l][j],  l > j
            const double bj = wm.sz[j], wj = wm.sw[j];
            double* row = wm.qu + upper_off(j, n) - j;
            if (on0 && i0 > j) { const double v = row[i0]; row[i0] = v + bj * (v * wj); }
            if (on1 && i1 > j) { const double v = row[i1]; row[i1] = v + bj * (v * wj); }
        }
        if (on0) wm.qu[upper_off(i0, n)] = nd0;
        if (on1) wm.qu[upper_off(i1, n)] = nd1;
        __syncwarp();
    }
    kappa *= r.delta;   // :122
    return status;
}

// ---------------------------------------------------------------------------------------------- kernel
template <bool DIAG, int MAXT>
__global__ void __launch_bounds__(MAXT) batch_stable_kernel(const BatchArgs a, int warps, int fstride, int per_warp_doubles, int lmi0) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = a.n, m = a.aux[0], tri = m * (m + 1) / 2;
    double* fp = smem;
    double* bp = fp + (long long)n * fstride;
    const long long base = lmi_stage_tables(a, smem, fstride, lmi0);   // F_k (packed lower, odd stride) and B, once per CTA
    double* ws = smem + base + (long long)warp * per_warp_doubles;
    WarpMem wm;
    wm.sx = ws; wm.sv = ws + kNMax; wm.sw = ws + 2 * kNMax; wm.sz = ws + 3 * kNMax;
    wm.sa = ws + 4 * kNMax;
    wm.st = wm.sa + ((tri + 1) & ~1);
    wm.swit = wm.st + m * m;
    wm.qu = wm.swit + ((m + 1) & ~1);
    LmiShared L{fp, bp, fstride, m, tri, n, lmi0};
    const CalcConsts cc = CalcConsts::make(n, a.use_parallel_cut);
    const int i0 = lane, i1 = lane + 32;
    const bool on0 = i0 < n, on1 = i1 < n;
    (void)warps;

    for (;;) {
        unsigned long long pid_u = 0;
        if (lane == 0) pid_u = atomicAdd(a.queue, 1ULL);
        pid_u = __shfl_sync(0xffffffffu, pid_u, 0);
        const long long pid = (long long)pid_u;
        if (pid >= a.count) break;

        double d0 = 0.0, d1 = 0.0;
        if (DIAG) {   // EllStable::new / new_with_scalar (src/ell_stable.rs:29-35)
            d0 = on0 ? (a.diag0 ? a.diag0[pid * a.diag0_stride + i0] : 1.0) : 0.0;
            d1 = on1 ? (a.diag0 ? a.diag0[pid * a.diag0_stride + i1] : 1.0) : 0.0;
        } else {      // new_with_matrix: diagonal + strict upper triangle are the state
            const double* mq = a.mq0 + pid * a.mq0_stride;
            for (int r = 0; r < n; ++r) {
                double* row = wm.qu + upper_off(r, n) - r;
                if (on0 && i0 >= r) row[i0] = mq[(long long)r * n + i0];
                if (on1 && i1 >= r) row[i1] = mq[(long long)r * n + i1];
            }
            __syncwarp();
        }
        double xc0 = on0 ? a.xc0[pid * a.xc0_stride + i0] : 0.0;
        double xc1 = on1 ? a.xc0[pid * a.xc0_stride + i1] : 0.0;
        double kappa = a.kappa0[pid * a.kappa0_stride];
        double tsq = 0.0;
        bool has_best = false;
        long long niter = 0;
        int exit_status = kExitMaxIt;

        for (; niter < a.max_iters; ++niter) {   // cutting_plane_feas, src/cutting_plane.rs:215-225
            wm.sx[lane] = xc0;
            if (i1 < kNMax) wm.sx[i1] = xc1;
            __syncwarp();
            double g0 = 0.0, g1 = 0.0, ep = 0.0;
            if (!lmi_assess(L, wm, lane, g0, g1, ep)) {
                has_best = true;
                exit_status = kExitFeasible;
                break;
            }
            const int status = stable_update<DIAG>(cc, kBias, wm, lane, n, g0, g1, single_cut(ep), d0, d1, xc0, xc1, kappa, tsq);
            if (status != kSuccess) { exit_status = status; break; }
            if (tsq < a.tolerance) { exit_status = kExitTol; break; }
        }
        if (lane == 0) {
            a.num_iters[pid] = niter;
            a.exit_status[pid] = exit_status;
            if (a.has_xbest) a.has_xbest[pid] = has_best ? 1 : 0;
            if (a.gamma) a.gamma[pid] = 0.0;
            if (a.tsq) a.tsq[pid] = tsq;
            if (a.kappa) a.kappa[pid] = kappa;
        }
        if (on0) {
            if (a.xbest && has_best) a.xbest[pid * n + i0] = xc0;
            if (a.xc) a.xc[pid * n + i0] = xc0;
        }
        if (on1) {
            if (a.xbest && has_best) a.xbest[pid * n + i1] = xc1;
            if (a.xc) a.xc[pid * n + i1] = xc1;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------- single EllStable handle
// One EllStable (n <= 64) driven call by call from the host (SearchSpace on EllStable, src/ell_stable.rs:128-210): one warp
// loads the packed upper triangle, runs the literal update_core recurrence (GENERAL variant above) and stores it back.
__global__ void __launch_bounds__(32) stable_single_kernel(double* __restrict__ qu_g, double* __restrict__ xc_g, double* __restrict__ scal,
                                                           int* __restrict__ status_out, const double* __restrict__ g, int n,
                                                           CalcConsts cc, int which, Cut cut) {
    ELL_STATIC_SHARED __align__(16) double s_qu[kNMax * (kNMax + 1) / 2];
    ELL_STATIC_SHARED __align__(16) double s_vec[3 * kNMax];
    const int lane = threadIdx.x;
    const int tri = n * (n + 1) / 2;
    for (int e = lane; e < tri; e += 32) s_qu[e] = qu_g[e];
    __syncwarp();
    WarpMem wm;
    wm.sx = nullptr; wm.sa = nullptr; wm.st = nullptr; wm.swit = nullptr;
    wm.sv = s_vec; wm.sw = s_vec + kNMax; wm.sz = s_vec + 2 * kNMax; wm.qu = s_qu;
    const int i0 = lane, i1 = lane + 32;
    const bool on0 = i0 < n, on1 = i1 < n;
    double g0 = on0 ? g[i0] : 0.0, g1 = on1 ? g[i1] : 0.0;
    double xc0 = on0 ? xc_g[i0] : 0.0, xc1 = on1 ? xc_g[i1] : 0.0;
    double kappa = scal[0], tsq = scal[1];
    double d0 = 0.0, d1 = 0.0;
    const int status = stable_update<false>(cc, which, wm, lane, n, g0, g1, cut, d0, d1, xc0, xc1, kappa, tsq);
    __syncwarp();
    if (status == kSuccess) {
        for (int e = lane; e < tri; e += 32) qu_g[e] = s_qu[e];
        if (on0) xc_g[i0] = xc0;
        if (on1) xc_g[i1] = xc1;
    }
    if (lane == 0) {
        scal[0] = kappa;   // unchanged on a failed cut
        scal[1] = tsq;     // tsq survives a failed cut (src/ell_stable.rs:85-90)
        *status_out = status;
    }
}

// ---------------------------------------------------------------------------------------------- single EllStable handle, 64 < n <= 2048
// The same update_core on the packed upper triangle in global memory, one CTA of 1024 threads, vectors in shared memory.
// The recurrences keep the reference's order (src/ell_stable.rs:61-121), which makes two of them serial by construction:
// the forward substitution needs w[j] final before column j is applied (n barriers), and row c of the back substitution
// subtracts its n-c-1 terms left to right starting with the most recently finished q[c+1] — a chain of n^2/2 dependent
// subtractions by one thread (10 clocks each).  It exists so that EllStable drops in at any n the reference user may have;
// the throughput path for EllStable is the batched kernel above.
using namespace stable_large;

__global__ void __launch_bounds__(kBigT) stable_large_kernel(double* __restrict__ qu, double* __restrict__ xc_g, double* __restrict__ scal,
                                                             int* __restrict__ status_out, const double* __restrict__ g, int n,
                                                             CalcConsts cc, int which, Cut cut) {
    extern __shared__ __align__(16) double sh[];
    __shared__ StableLargeShared ss;
    stable_large_body(qu, xc_g, scal, status_out, g, n, cc, which, cut, sh, ss, (int)threadIdx.x);
}

}  // namespace

#ifdef __CUDACC__   // host launchers: left out when tests/cpp/emu_lmi.cpp compiles the kernels above for the host
int launch_stable_single(double* qu, double* xc, double* scal, int* status, const double* g, int n, int use_parallel_cut, int which,
                         const Cut& cut, cudaStream_t s) {
    const CalcConsts cc = CalcConsts::make(n, use_parallel_cut);
    if (n <= kNMax) {
        stable_single_kernel<<<1, 32, 0, s>>>(qu, xc, scal, status, g, n, cc, which, cut);
    } else {
        if (n > kBigNMax) { set_error("a single EllStable handle supports n <= 2048"); return ELLCUDA_E_ARG; }
        const size_t smem = (size_t)7 * n * sizeof(double);
        ELL_CUDA_TRY(cudaFuncSetAttribute(stable_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        stable_large_kernel<<<1, kBigT, smem, s>>>(qu, xc, scal, status, g, n, cc, which, cut);
    }
    ELL_CUDA_TRY(cudaGetLastError());
    return ELLCUDA_OK;
}

int launch_batch_stable(const BatchArgs& a, int oracle_id, cudaStream_t s, int sm_count, long long* launches) {
    if (oracle_id != ELL_ORACLE_LMI && oracle_id != ELL_ORACLE_LMI_OLD && oracle_id != ELL_ORACLE_LMI0) {
        set_error("this oracle has no device implementation yet");
        return ELLCUDA_E_ARG;
    }
    // LMI_OLD: the reference builds all m*m elements (k outermost) and then factorizes; every element undergoes the same
    // operations in the same order as in LMIOracle's lazy closure, so the lazy kernel is its exact equivalent.
    const int lmi0 = oracle_id == ELL_ORACLE_LMI0;
    if (a.kind == ELL_KIND_ELL) return launch_batch_ell64(a, oracle_id, s, sm_count, launches);   // the same oracles on Ell (M in TMEM)
    if (a.mode != ELL_MODE_FEAS) {
        set_error("the LMI device oracles run under cutting_plane_feas");
        return ELLCUDA_E_ARG;
    }
    const int n = a.n, m = a.aux[0];
    if (n < 1 || n > kNMax) { set_error("LMI kernel needs 1 <= n <= 64"); return ELLCUDA_E_ARG; }
    if (m < 1 || m > 32) { set_error("LMI kernel needs 1 <= m <= 32"); return ELLCUDA_E_ARG; }
    if (!a.table || a.table_len < (long long)(n + (lmi0 ? 0 : 1)) * m * m) {
        set_error(lmi0 ? "LMI0 needs table = F_0..F_{n-1}" : "LMI needs table = F_0..F_{n-1}, B");
        return ELLCUDA_E_ARG;
    }
    const int tri = m * (m + 1) / 2;
    const int fstride = tri | 1;
    const bool diag = a.mq0 == nullptr;
    int per_warp = 4 * kNMax + ((tri + 1) & ~1) + m * m + ((m + 1) & ~1);
    if (!diag) per_warp += (n * (n + 1) / 2 + 1) & ~1;
    per_warp = (per_warp + 1) & ~1;
    const size_t fixed = (((size_t)n * fstride + tri + 1) & ~(size_t)1) * 8;
    const size_t budget = 227 * 1024;
    if (fixed + (size_t)per_warp * 8 > budget) { set_error("F table too large for shared memory"); return ELLCUDA_E_ARG; }
    int warps = (int)((budget - fixed) / ((size_t)per_warp * 8));
    // measured (n=64, m=16, diagonal start, 32768 problems): 8 warps 38.9, 16 58.9, 20 63.3, 24 65.4, 28 67.0 M iter/s —
    // the serial LDLT division chain wants every warp the register file can hold
    int cap = 28;
    cap = ELL_ENV_INT("ELLCUDA_LMI_WARPS", 28);   // performance experiments (read once per process)
    if (cap < 1) cap = 1;
    if (cap > 28) cap = 28;
    if (!diag && cap > 16) cap = 16;   // the GENERAL variant is compiled for at most 512 threads
    if (warps > cap) warps = cap;
    const size_t smem = fixed + (size_t)warps * per_warp * 8;
    long long blocks = sm_count;
    const long long need = (a.count + warps - 1) / warps;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    auto launch = [&](auto kern) -> int {
        ELL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)blocks, warps * 32, smem, s>>>(a, warps, fstride, per_warp, lmi0);
        return ELLCUDA_OK;
    };
    int rc;
    if (diag) rc = warps > 20 ? launch(batch_stable_kernel<true, 896>) : warps > 16 ? launch(batch_stable_kernel<true, 640>)
                                                                                     : launch(batch_stable_kernel<true, 512>);
    else rc = launch(batch_stable_kernel<false, 512>);
    if (rc) return rc;
    ELL_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += 1;
    return ELLCUDA_OK;
}

#endif  // __CUDACC__

}  // namespace ellcuda
