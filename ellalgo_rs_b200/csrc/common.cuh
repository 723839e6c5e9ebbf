// common.cuh — kernel argument records and small helpers shared by the .cu files of libellcuda.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <string>

#include <nvtx3/nvToolsExt.h>

#include "../../include/ellcuda.h"
#include "ell_calc.cuh"

namespace ellcuda {

// Device-side picture of ell_batch_desc + ell_batch_result (all pointers are device pointers).
struct BatchArgs {
    int n;
    int mode;
    int kind;
    int no_defer_trick;
    int use_parallel_cut;
    int params_width;
    long long count;
    long long max_iters;
    double tolerance;
    const double* kappa0; long long kappa0_stride;
    const double* xc0;    long long xc0_stride;
    const double* diag0;  long long diag0_stride;
    const double* mq0;    long long mq0_stride;
    const double* gamma0; long long gamma0_stride;
    const double* params; long long params_stride;
    const double* table;  long long table_len;
    int aux[8];
    // results (may be null except num_iters / exit_status which the host always provides)
    double* xbest;
    unsigned char* has_xbest;
    long long* num_iters;
    int* exit_status;
    double* gamma;
    double* tsq;
    double* xc;
    double* kappa;
    unsigned long long* queue;  // dynamic work counter (persistent kernels)
    const double* table_t;      // launcher-made scratch: `table` transposed (batch_ell64.cu / batch_ellcta.cu, lowpass with n > 32), else null
};

// NVTX range for the scope of a host-side call (tracing); a no-op unless a profiler is attached
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// integer tuning switch from the environment, read once per process (performance experiments only)
inline int env_int_once(const char* name, int dflt, int* cache, bool* have) {
    if (!*have) {
        const char* e = getenv(name);
        *cache = e ? atoi(e) : dflt;
        *have = true;
    }
    return *cache;
}
#define ELL_ENV_INT(name, dflt) ([]() -> int { static int v = 0; static bool have = false; return ::ellcuda::env_int_once(name, dflt, &v, &have); }())

// `__shared__` for a kernel's static shared variables.  nvcc sees exactly `__shared__`; the host emulation of the kernels
// (tests/cpp/emu_*.cpp, which compile these sources with g++ and run one emulated CTA at a time) sees `static`.
#ifdef __CUDACC__
#define ELL_STATIC_SHARED __shared__
#else
#define ELL_STATIC_SHARED static
#endif

void set_error(const std::string& msg);
void keep_mempool(int dev);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define ELL_CUDA_TRY(expr)                                                         \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) return ::ellcuda::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

// makes `dev` current for the scope of a C-ABI call and restores the caller's device afterwards (the library must
// not change the process' current device behind torch's back)
struct DevGuard {
    int prev = -1;
    cudaError_t err;
    explicit DevGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        err = cudaSetDevice(dev);
    }
    ~DevGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define ELL_DEVICE_SCOPE(dev)           \
    ::ellcuda::DevGuard _dev_guard(dev); \
    ELL_CUDA_TRY(_dev_guard.err)

// launchers implemented per kernel family; each returns ELLCUDA_* and adds to *launches
int launch_batch_small(const BatchArgs& a, int oracle_id, cudaStream_t s, int sm_count, long long* launches);
int launch_batch_warp(const BatchArgs& a, int oracle_id, cudaStream_t s, int sm_count, long long* launches);
int launch_batch_lowpass32(const BatchArgs& a, cudaStream_t s, int sm_count, long long* launches);
int launch_math_probe(int which, const double* x_dev, double* y_dev, long long n, cudaStream_t s);
int launch_stable_single(double* qu_packed_upper, double* xc, double* scal_kappa_tsq, int* status, const double* g, int n,
                         int use_parallel_cut, int which, const Cut& cut, cudaStream_t s);
int launch_batch_stable(const BatchArgs& a, int oracle_id, cudaStream_t s, int sm_count, long long* launches);
int launch_batch_ell64(const BatchArgs& a, int oracle_id, cudaStream_t s, int sm_count, long long* launches);
int launch_batch_ellcta(const BatchArgs& a, int oracle_id, cudaStream_t s, int sm_count, long long* launches);
int launch_transpose_table(const double* in, double* out, int rows, int cols, cudaStream_t s);

}  // namespace ellcuda
