// batch_ellcta.cu — batched cutting-plane loops on `Ell` (src/ell.rs) for 32 < n <= 224: ONE CTA PER PROBLEM, persistent,
// with the shape matrix M in SHARED MEMORY as a packed lower triangle.
//
// batch_warp.cu (n <= 32, rows in registers) and batch_ell64.cu (n <= 64, rows in tensor memory) give a problem to one
// warp.  Above n = 64 a row no longer fits a lane pair, and the reference's own stress test runs LowpassOracle with 128 taps
// (tests/stress_tests.rs:6-24), so here thread t of the CTA owns row t.  (It is also the faster kernel for most of
// 32 < n <= 64 — see the routing in batch_warp.cu — which leaves the tensor-memory kernel the LMI oracles on `Ell`.)
//
// The reference computes the lower triangle of the rank-1 update and mirrors it (src/ell.rs:117-128), so after the first
// successful cut M is bitwise symmetric and only P[i(i+1)/2 + j] = M[i][j], j <= i, is kept: 66 KB at n = 128 (3 resident
// CTAs per SM), 201 KB at n = 224, for all of the problem's iterations — an update touches no HBM at all.  Row t of M.g
// reads P[T(t)+j] for j <= t and P[T(j)+t] above the diagonal: the triangular numbers T(t) of 16 consecutive t are distinct
// mod 16, so both forms are free of bank conflicts for 8-byte words.  A non-symmetric new_with_matrix start
// (src/ell.rs:31-41) is multiplied from the caller's rows in global memory for the one product that can see it
// (iteration 0: a failed cut ends the loop, a successful one mirrors the lower triangle).  The serial sums the whole CTA
// needs (omega, the bowl's f) and EllCalc run in warp 0 only and are broadcast, which leaves the FP64 pipe to the other
// resident CTAs.
//
//   cutting_plane_optim                  src/cutting_plane.rs:286-313
//   Ell::update_core                     src/ell.rs:97-137 (grad_t, omega, xc, lower-triangle rank-1 + mirror, kappa)
//   oracles: quadratic bowl (tests/integration_test.rs:90-105), LowpassOracle (src/oracles/lowpass_oracle.rs:7-150; one
//            spectrum row per thread and round, the table read transposed from L2)
// Every sum keeps the reference's order (row dots j ascending from 0.0, omega k ascending), -fmad=false: results are
// bit-identical to the Rust path.
#include "ellcta_device.cuh"

namespace ellcuda {
namespace {

using namespace ellcta;

template <class Oracle>
__global__ void __launch_bounds__(kCtaNMax, 1) batch_ellcta_kernel(const BatchArgs a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ CtaShared sh;
    ellcta_body<Oracle>(a, smem, sh, (int)threadIdx.x, (int)blockDim.x);
}

template <class Oracle>
int launch_k(const BatchArgs& a, cudaStream_t s, int sm_count) {
    const int n = a.n, nt = ((n + 31) / 32) * 32;
    const size_t smem = ((((size_t)n * (n + 1) / 2 + 1) & ~(size_t)1) + 4 * (size_t)nt) * 8;
    if (smem > 227 * 1024) { set_error("matrix does not fit in shared memory"); return ELLCUDA_E_ARG; }
    auto kern = batch_ellcta_kernel<Oracle>;
    ELL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    ELL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nt, smem));
    if (per_sm < 1) per_sm = 1;
    long long blocks = (long long)sm_count * per_sm;   // persistent: one wave exactly
    if (blocks > a.count) blocks = a.count;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, nt, smem, s>>>(a);
    ELL_CUDA_TRY(cudaGetLastError());
    return ELLCUDA_OK;
}

}  // namespace

// Ell (KIND_ELL) with n <= 224 under cutting_plane_optim: bowl / lowpass
int launch_batch_ellcta(const BatchArgs& a_in, int oracle_id, cudaStream_t s, int sm_count, long long* launches) {
    BatchArgs a = a_in;
    if (a.kind != ELL_KIND_ELL || a.mode != ELL_MODE_OPTIM) {
        set_error("CTA-per-problem kernel: Ell + cutting_plane_optim only");
        return ELLCUDA_E_ARG;
    }
    if (a.n < 1 || a.n > kCtaNMax) { set_error("batched Ell needs 1 <= n <= 224"); return ELLCUDA_E_ARG; }
    int rc;
    if (oracle_id == ELL_ORACLE_BOWL) {
        if (!a.params || a.params_width < a.n) { set_error("BOWL needs params[n] = target"); return ELLCUDA_E_ARG; }
        rc = launch_k<OrBowlCta>(a, s, sm_count);
    } else if (oracle_id == ELL_ORACLE_LOWPASS) {
        if (!a.params || a.params_width < 5 || !a.table || a.aux[0] != 15 * a.n || a.table_len < (long long)a.aux[0] * a.n) {
            set_error("LOWPASS needs params[5], the spectrum table and aux[0] = 15*n");
            return ELLCUDA_E_ARG;
        }
        double* tt = nullptr;   // the spectrum transposed: threads evaluating consecutive rows then read consecutive addresses
        ELL_CUDA_TRY(cudaMallocAsync((void**)&tt, (size_t)a.aux[0] * a.n * 8, s));
        rc = launch_transpose_table(a.table, tt, a.aux[0], a.n, s);
        if (rc == ELLCUDA_OK) {
            a.table_t = tt;
            if (launches) *launches += 1;
            rc = launch_k<OrLowpassCta>(a, s, sm_count);
        }
        cudaFreeAsync(tt, s);
    } else {
        set_error("oracle not available on the batched Ell (n > 64) kernel");
        return ELLCUDA_E_ARG;
    }
    if (rc == ELLCUDA_OK && launches) *launches += 1;
    return rc;
}

}  // namespace ellcuda
