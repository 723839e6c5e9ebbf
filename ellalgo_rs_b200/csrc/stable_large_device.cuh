// stable_large_device.cuh — device code of the single-handle EllStable update for 64 < n <= 2048 (batch_stable.cu holds the
// __global__ entry and the launcher).  In a header so that tests/cpp/emu_stable_large.cpp can compile the same code for the
// host (one OS thread per CUDA thread, pthread barriers for __syncthreads) and run it against the checker and under
// ThreadSanitizer.
#pragma once
#include "common.cuh"

namespace ellcuda {
namespace stable_large {

#ifndef ELL_STABLE_LARGE_THREADS
#define ELL_STABLE_LARGE_THREADS 1024   // the host emulation under ThreadSanitizer builds with 128 (it tracks <= 256 running threads)
#endif
constexpr int kBigT = ELL_STABLE_LARGE_THREADS;
constexpr int kBigNMax = 2048;
__device__ __forceinline__ long long upper_off_ll(long long r, long long n) { return r * n - r * (r - 1) / 2; }

struct StableLargeShared {   // the CTA's static shared memory
    double omega, roo, delta;
    int status;
};

// the whole kernel as seen by thread `tid` of the CTA of kBigT threads; `sh` = its dynamic shared memory (7n doubles)
__device__ __forceinline__ void stable_large_body(double* __restrict__ qu, double* __restrict__ xc_g, double* __restrict__ scal,
                                                  int* __restrict__ status_out, const double* __restrict__ g, int n,
                                                  const CalcConsts& cc, int which, const Cut& cut, double* sh, StableLargeShared& ss,
                                                  const int tid) {
    double* w = sh;               // inv(L) g
    double* z = w + n;            // w .* d
    double* q = z + n;            // mq * grad (back substitution)
    double* gg = q + n;           // z .* w
    double* t = gg + n;           // terms of the current back-substitution row; later beta2
    double* oldt_a = t + n;       // oldt before adding gg[j]
    double* temp_a = oldt_a + n;  // oldt + gg[j]
    double& s_omega = ss.omega;
    double& s_roo = ss.roo;
    double& s_delta = ss.delta;
    int& s_status = ss.status;
    for (int i = tid; i < n; i += kBigT) w[i] = g[i];
    // --- inv(L)*grad (:60-69): column j is applied once w[j] is final
    for (int j = 0; j + 1 < n; ++j) {
        __syncthreads();
        const double wj = w[j];
        const double* row = qu + upper_off_ll(j, n) - j;      // row[c] = Q[j][c], c >= j
        for (int i = j + 1 + tid; i < n; i += kBigT) w[i] -= row[i] * wj;
    }
    __syncthreads();
    // --- inv(D)*inv(L)*grad and omega (:72-83)
    for (int i = tid; i < n; i += kBigT) {
        const double d = qu[upper_off_ll(i, n)];
        const double zi = w[i] * d;
        z[i] = zi;
        q[i] = zi;
        gg[i] = zi * w[i];
    }
    __syncthreads();
    if (tid == 0) {
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc += gg[k];
        const double tsq = scal[0] * acc;
        Rsd r;
        const int status = ell_calc(cc, which, cut, tsq, r);
        scal[1] = tsq;                                        // tsq survives a failed cut (:85-90)
        *status_out = status;
        s_status = status;
        s_omega = acc;
        if (status == kSuccess) {
            s_roo = r.rho / acc;
            s_delta = r.delta;
            const double mu = r.sigma / (1.0 - r.sigma);      // :107-108
            double oldt = acc / mu;
            for (int j = 0; j < n; ++j) {                     // temp = oldt + gg[j]; oldt = temp (:109-121)
                const double temp = oldt + gg[j];
                oldt_a[j] = oldt;
                temp_a[j] = temp;
                oldt = temp;
            }
        }
    }
    __syncthreads();
    if (s_status != kSuccess) return;                         // state unchanged except tsq
    // --- mq*grad = inv(L')*inv(D)*inv(L)*grad (:93-98): q[c] = z[c] - sum_{j>c, ascending} (Q[c][j]*w[c])*q[j]
    for (int c = n - 2; c >= 0; --c) {
        const double wc = w[c];
        const double* row = qu + upper_off_ll(c, n) - c;
        for (int j = c + 1 + tid; j < n; j += kBigT) t[j] = (row[j] * wc) * q[j];
        __syncthreads();
        if (tid == 0) {
            double qc = q[c];
            for (int j = c + 1; j < n; ++j) qc -= t[j];
            q[c] = qc;
        }
        __syncthreads();
    }
    // --- xc (:101-104), beta2 and the new diagonal (:110-120)
    const double roo = s_roo;
    for (int i = tid; i < n; i += kBigT) {
        xc_g[i] -= roo * q[i];
        t[i] = z[i] / temp_a[i];                              // beta2 = inv_md_inv_ml_g[j] / temp
    }
    __syncthreads();
    // --- rank-one update of the factor: mq[j][l] += beta2 * mq[l][j], l > j, with the scratch entry mq[l][j] = Q[j][l]*w[j]
    for (int j = 0; j < n; ++j) {
        double* row = qu + upper_off_ll(j, n) - j;
        const double bj = t[j], wj = w[j];
        for (int l = j + 1 + tid; l < n; l += kBigT) {
            const double v = row[l];
            row[l] = v + bj * (v * wj);
        }
        if (tid == 0) row[j] = row[j] * (oldt_a[j] / temp_a[j]);   // mq[j][j] *= oldt / temp
    }
    if (tid == 0) scal[0] = scal[0] * s_delta;                // kappa *= delta (:122)
}


}  // namespace stable_large
}  // namespace ellcuda
