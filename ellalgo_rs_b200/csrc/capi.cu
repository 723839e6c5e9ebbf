// capi.cu — C ABI of libellcuda.so, batched mode + misc (include/ellcuda.h).
// Host-side glue only: validates the descriptor, moves host buffers to the device when asked to, picks the kernel
// family, and reports errors.  There is no CPU implementation of anything here: without a usable sm_100 device
// every compute entry point returns ELLCUDA_E_CUDA.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace ellcuda {

static thread_local std::string g_err = "";
static thread_local long long g_last_launches = 0;

void set_error(const std::string& msg) { g_err = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    g_err = buf;
    return e == cudaErrorMemoryAllocation ? ELLCUDA_E_OOM : ELLCUDA_E_CUDA;
}

// cudaMallocAsync's default pool gives memory back to the driver at every synchronisation unless told otherwise;
// the scratch buffers of consecutive calls would then be re-created each time (hundreds of microseconds).
void keep_mempool(int dev) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ULL;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
}

namespace {

struct DeviceInfo {
    bool ok = false;
    int sm_count = 0;
    int major = 0;
    cudaStream_t stream = nullptr;
};
std::mutex g_dev_mu;
std::vector<DeviceInfo> g_devs;

int device_info(int dev, DeviceInfo** out) {
    std::lock_guard<std::mutex> lk(g_dev_mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (libellcuda has no CPU fallback)");
        return ELLCUDA_E_CUDA;
    }
    if (dev < 0 || dev >= count) { set_error("device ordinal out of range"); return ELLCUDA_E_ARG; }
    if ((int)g_devs.size() < count) g_devs.resize(count);
    DeviceInfo& d = g_devs[dev];
    if (!d.ok) {
        cudaDeviceProp p;
        ELL_CUDA_TRY(cudaGetDeviceProperties(&p, dev));
        if (p.major < 10) { set_error("libellcuda is built for sm_100a (B200) only"); return ELLCUDA_E_CUDA; }
        ELL_DEVICE_SCOPE(dev);
        ELL_CUDA_TRY(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
        keep_mempool(dev);
        d.sm_count = p.multiProcessorCount;
        d.major = p.major;
        d.ok = true;
    }
    *out = &d;
    return ELLCUDA_OK;
}

// temp device allocations released at scope exit
struct Scratch {
    cudaStream_t s;
    std::vector<void*> ptrs;
    explicit Scratch(cudaStream_t st) : s(st) {}
    ~Scratch() { for (void* p : ptrs) cudaFreeAsync(p, s); }
    int alloc(void** out, size_t bytes) {
        if (bytes == 0) bytes = 8;
        cudaError_t e = cudaMallocAsync(out, bytes, s);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync", __FILE__, __LINE__);
        ptrs.push_back(*out);
        return ELLCUDA_OK;
    }
};

int validate(const ell_batch_desc* d, const ell_batch_result* r) {
    if (!d || !r) { set_error("null descriptor"); return ELLCUDA_E_ARG; }
    if (d->n <= 0) { set_error("n must be positive"); return ELLCUDA_E_ARG; }
    if (d->count < 0) { set_error("count must be non-negative"); return ELLCUDA_E_ARG; }
    if (d->max_iters < 0) { set_error("max_iters must be non-negative"); return ELLCUDA_E_ARG; }
    if (!d->kappa0 || !d->xc0) { set_error("kappa0 and xc0 are required"); return ELLCUDA_E_ARG; }
    if (d->kind != ELL_KIND_ELL && d->kind != ELL_KIND_ELLSTABLE) { set_error("bad kind"); return ELLCUDA_E_ARG; }
    if (d->mode < ELL_MODE_OPTIM || d->mode > ELL_MODE_BSEARCH) { set_error("bad mode"); return ELLCUDA_E_ARG; }
    if (d->oracle_id < 0 || d->oracle_id >= ELL_ORACLE_COUNT) { set_error("bad oracle_id"); return ELLCUDA_E_ARG; }
    if (!r->num_iters || !r->exit_status) { set_error("num_iters and exit_status outputs are required"); return ELLCUDA_E_ARG; }
    if (d->mode != ELL_MODE_FEAS && !d->gamma0) { set_error("gamma0 is required in optim modes"); return ELLCUDA_E_ARG; }
    if (d->params && d->params_width <= 0) { set_error("params_width must be positive"); return ELLCUDA_E_ARG; }
    return ELLCUDA_OK;
}

}  // namespace

static int batch_solve_impl(const ell_batch_desc* d, const ell_batch_result* r) {
    g_last_launches = 0;
    NvtxRange nvtx_range("ellcuda_batch_solve");
    int rc = validate(d, r);
    if (rc) return rc;
    DeviceInfo* di = nullptr;
    rc = device_info(d->device, &di);
    if (rc) return rc;
    ELL_DEVICE_SCOPE(d->device);
    cudaStream_t s = d->stream ? (cudaStream_t)d->stream : di->stream;
    if (d->count == 0) return ELLCUDA_OK;
    // mq0 need not be symmetric (Ell::new_with_matrix takes any Arr, src/ell.rs:31-41): the kernels multiply by the rows as
    // given until the first successful cut and then rewrite the upper triangle from the lower one, as src/ell.rs:117-128 does

    if ((d->oracle_id == ELL_ORACLE_LMI || d->oracle_id == ELL_ORACLE_LMI_OLD || d->oracle_id == ELL_ORACLE_LMI0) && !d->on_device &&
        d->table) {
        // the device keeps F_k as packed lower triangles; sym_quad (ldlt_mgr.rs:115-124) reads both triangles, so the
        // F_k must be symmetric — which an LMI requires anyway
        const long long m = d->aux[0];
        if (m >= 1 && d->table_len >= (long long)d->n * m * m)
            for (long long k = 0; k < d->n; ++k)
                for (long long i = 0; i < m; ++i)
                    for (long long j = 0; j < i; ++j)
                        if (!(d->table[k * m * m + i * m + j] == d->table[k * m * m + j * m + i])) {
                            set_error("LMI: every F_k must be symmetric");
                            return ELLCUDA_E_ARG;
                        }
    }

    Scratch scratch(s);
    BatchArgs a;
    memset(&a, 0, sizeof a);
    a.n = d->n; a.mode = d->mode; a.kind = d->kind;
    a.no_defer_trick = d->no_defer_trick; a.use_parallel_cut = d->use_parallel_cut;
    a.params_width = d->params_width;
    a.count = d->count; a.max_iters = d->max_iters; a.tolerance = d->tolerance;
    a.kappa0_stride = d->kappa0_stride; a.xc0_stride = d->xc0_stride; a.diag0_stride = d->diag0_stride;
    a.mq0_stride = d->mq0_stride; a.gamma0_stride = d->gamma0_stride; a.params_stride = d->params_stride;
    a.table_len = d->table_len;
    for (int k = 0; k < 8; ++k) a.aux[k] = d->aux[k];

    const long long n = d->n, cnt = d->count;
    auto span = [&](long long stride, long long width) { return (size_t)((stride ? (cnt - 1) * stride : 0) + width); };

    struct Out { void* host; void* dev; size_t bytes; };
    std::vector<Out> outs;

    if (d->on_device) {
        a.kappa0 = d->kappa0; a.xc0 = d->xc0; a.diag0 = d->diag0; a.mq0 = d->mq0; a.gamma0 = d->gamma0;
        a.params = d->params; a.table = d->table;
        a.xbest = r->xbest; a.has_xbest = r->has_xbest; a.num_iters = (long long*)r->num_iters;
        a.exit_status = r->exit_status; a.gamma = r->gamma; a.tsq = r->tsq; a.xc = r->xc; a.kappa = r->kappa;
    } else {
        auto up = [&](const double* host, size_t count_d, const double** dev) -> int {
            *dev = nullptr;
            if (!host) return ELLCUDA_OK;
            void* p = nullptr;
            int e = scratch.alloc(&p, count_d * sizeof(double));
            if (e) return e;
            ELL_CUDA_TRY(cudaMemcpyAsync(p, host, count_d * sizeof(double), cudaMemcpyHostToDevice, s));
            *dev = (const double*)p;
            return ELLCUDA_OK;
        };
        if ((rc = up(d->kappa0, span(d->kappa0_stride, 1), &a.kappa0))) return rc;
        if ((rc = up(d->xc0, span(d->xc0_stride, n), &a.xc0))) return rc;
        if ((rc = up(d->diag0, span(d->diag0_stride, n), &a.diag0))) return rc;
        if ((rc = up(d->mq0, span(d->mq0_stride, n * n), &a.mq0))) return rc;
        if (d->mode != ELL_MODE_FEAS && (rc = up(d->gamma0, span(d->gamma0_stride, 1), &a.gamma0))) return rc;
        if ((rc = up(d->params, span(d->params_stride, d->params_width), &a.params))) return rc;
        if ((rc = up(d->table, (size_t)d->table_len, &a.table))) return rc;
        auto down = [&](void* host, size_t bytes, void** dev, bool zero) -> int {
            *dev = nullptr;
            if (!host) return ELLCUDA_OK;
            int e = scratch.alloc(dev, bytes);
            if (e) return e;
            if (zero) ELL_CUDA_TRY(cudaMemsetAsync(*dev, 0, bytes, s));
            outs.push_back({host, *dev, bytes});
            return ELLCUDA_OK;
        };
        if ((rc = down(r->xbest, (size_t)cnt * n * 8, (void**)&a.xbest, true))) return rc;
        if ((rc = down(r->has_xbest, (size_t)cnt, (void**)&a.has_xbest, false))) return rc;
        if ((rc = down(r->num_iters, (size_t)cnt * 8, (void**)&a.num_iters, false))) return rc;
        if ((rc = down(r->exit_status, (size_t)cnt * 4, (void**)&a.exit_status, false))) return rc;
        if ((rc = down(r->gamma, (size_t)cnt * 8, (void**)&a.gamma, false))) return rc;
        if ((rc = down(r->tsq, (size_t)cnt * 8, (void**)&a.tsq, false))) return rc;
        if ((rc = down(r->xc, (size_t)cnt * n * 8, (void**)&a.xc, false))) return rc;
        if ((rc = down(r->kappa, (size_t)cnt * 8, (void**)&a.kappa, false))) return rc;
    }
    void* q = nullptr;
    if ((rc = scratch.alloc(&q, sizeof(unsigned long long)))) return rc;
    ELL_CUDA_TRY(cudaMemsetAsync(q, 0, sizeof(unsigned long long), s));
    a.queue = (unsigned long long*)q;

    switch (d->oracle_id) {
        case ELL_ORACLE_BOWL:
        case ELL_ORACLE_LOWPASS:
            rc = launch_batch_warp(a, d->oracle_id, s, di->sm_count, &g_last_launches);
            break;
        case ELL_ORACLE_LMI:
        case ELL_ORACLE_LMI_OLD:
        case ELL_ORACLE_LMI0:
            rc = launch_batch_stable(a, d->oracle_id, s, di->sm_count, &g_last_launches);
            break;
        default:
            rc = launch_batch_small(a, d->oracle_id, s, di->sm_count, &g_last_launches);
            break;
    }
    if (rc) return rc;

    if (!d->on_device) {
        for (const Out& o : outs) ELL_CUDA_TRY(cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, s));
        ELL_CUDA_TRY(cudaStreamSynchronize(s));
    }
    return ELLCUDA_OK;
}

}  // namespace ellcuda

extern "C" {

int ellcuda_batch_solve(const ell_batch_desc* desc, const ell_batch_result* res) {
    return ellcuda::batch_solve_impl(desc, res);
}

// SURVEY.md §8b `ndev, devs` / §8e: independent problems, contiguous blocks, one host thread and one device per block.
int ellcuda_batch_solve_multi(const ell_batch_desc* desc, const ell_batch_result* res, int32_t ndev, const int32_t* devs) {
    using namespace ellcuda;
    if (!desc || !res || !devs) { set_error("null argument"); return ELLCUDA_E_ARG; }
    if (ndev < 1 || ndev > 64) { set_error("ndev must be in [1, 64]"); return ELLCUDA_E_ARG; }
    if (desc->on_device) { set_error("ellcuda_batch_solve_multi takes host buffers"); return ELLCUDA_E_ARG; }
    if (desc->count < 0) { set_error("count must be non-negative"); return ELLCUDA_E_ARG; }
    const int64_t cnt = desc->count, n = desc->n;
    std::vector<int> rcs((size_t)ndev, ELLCUDA_OK);
    std::vector<std::string> errs((size_t)ndev);
    std::vector<long long> launches((size_t)ndev, 0);
    auto work = [&](int r) {
        const int64_t base = cnt / ndev, rem = cnt % ndev;
        const int64_t lo = r * base + (r < rem ? r : rem), hi = lo + base + (r < rem ? 1 : 0);
        ell_batch_desc d = *desc;
        ell_batch_result o = *res;
        d.count = hi - lo;
        d.device = devs[r];
        d.stream = nullptr;
        auto adv = [&](const double*& p, int64_t stride) { if (p) p += lo * stride; };
        adv(d.kappa0, d.kappa0_stride); adv(d.xc0, d.xc0_stride); adv(d.diag0, d.diag0_stride); adv(d.mq0, d.mq0_stride);
        adv(d.gamma0, d.gamma0_stride); adv(d.params, d.params_stride);
        if (o.xbest) o.xbest += lo * n;
        if (o.has_xbest) o.has_xbest += lo;
        if (o.num_iters) o.num_iters += lo;
        if (o.exit_status) o.exit_status += lo;
        if (o.gamma) o.gamma += lo;
        if (o.tsq) o.tsq += lo;
        if (o.xc) o.xc += lo * n;
        if (o.kappa) o.kappa += lo;
        rcs[r] = batch_solve_impl(&d, &o);
        if (rcs[r]) errs[r] = g_err;          // thread-local on the worker: hand it to the caller
        launches[r] = g_last_launches;
    };
    if (ndev == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int r = 0; r < ndev; ++r) th.emplace_back(work, r);
        for (auto& t : th) t.join();
    }
    g_last_launches = 0;
    for (int r = 0; r < ndev; ++r) g_last_launches += launches[r];
    for (int r = 0; r < ndev; ++r)
        if (rcs[r]) { set_error("device " + std::to_string(devs[r]) + ": " + errs[r]); return rcs[r]; }
    return ELLCUDA_OK;
}

int64_t ellcuda_last_launch_count(void) { return ellcuda::g_last_launches; }

// LowpassOracle::new's table (src/oracles/lowpass_oracle.rs:24-33; linspace src/arr.rs:506-520).  Host code on
// purpose: glibc's cos is what the Rust constructor calls, so the device reads bit-identical coefficients.
int ellcuda_lowpass_table(int32_t n, double* out) {
    if (n <= 0 || !out) { ellcuda::set_error("bad arguments"); return ELLCUDA_E_ARG; }
    const long long mdim = 15LL * n;
    const double pi = 3.14159265358979323846264338327950288;
    const double step = (mdim > 1) ? (pi - 0.0) / (double)(mdim - 1) : 0.0;
    for (long long i = 0; i < mdim; ++i) {
        const double w = (mdim == 1) ? 0.0 : 0.0 + step * (double)i;
        out[i * n] = 1.0;
        for (long long j = 1; j < n; ++j) out[i * n + j] = 2.0 * std::cos(w * (double)j);
    }
    return ELLCUDA_OK;
}

int32_t ellcuda_lowpass_band_edge(int32_t n, double w) {
    const long long mdim = 15LL * n;
    return (int32_t)std::floor(w * (double)(mdim - 1)) + 1;
}

// Diagnostic: y[i] = exp(x[i]) (which = 0) or ln(x[i]) (which = 1) with the DEVICE routines the oracles use
// (glibc_math.cuh), host buffers in and out.  Lets a test prove they equal the host libm bit for bit.
int ellcuda_math_probe(int32_t which, const double* x_host, double* y_host, int64_t n, int32_t device) {
    using namespace ellcuda;
    if (!x_host || !y_host || n < 0 || which < 0 || which > 1) { set_error("bad arguments"); return ELLCUDA_E_ARG; }
    DeviceInfo* di = nullptr;
    int rc = device_info(device, &di);
    if (rc) return rc;
    ELL_DEVICE_SCOPE(device);
    if (n == 0) return ELLCUDA_OK;
    cudaStream_t s = di->stream;
    Scratch scratch(s);
    void *dx = nullptr, *dy = nullptr;
    if ((rc = scratch.alloc(&dx, (size_t)n * 8)) || (rc = scratch.alloc(&dy, (size_t)n * 8))) return rc;
    ELL_CUDA_TRY(cudaMemcpyAsync(dx, x_host, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    if ((rc = launch_math_probe(which, (const double*)dx, (double*)dy, n, s))) return rc;
    ELL_CUDA_TRY(cudaMemcpyAsync(y_host, dy, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    ELL_CUDA_TRY(cudaStreamSynchronize(s));
    return ELLCUDA_OK;
}

const char* ellcuda_last_error(void) { return ellcuda::g_err.c_str(); }
int ellcuda_version(void) { return 100; }
int ellcuda_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < count; ++i) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major >= 10) ++ok;
    }
    return ok;
}

}  // extern "C"
