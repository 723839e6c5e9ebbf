// emu_batch_warp.cpp — the warp-per-problem batched `Ell` kernels of ellalgo_rs_b200/csrc/batch_warp.cu (n <= 32, M's rows in
// registers: the quadratic bowl incl. the 2^20-problem shape of bench.py's `bowl_1m`, LowpassOracle for n < 32 and for
// new_with_matrix starts) compiled for the host with cuda_emu.h.  The .cu file is included as it is (its host launchers sit
// behind #ifdef __CUDACC__); the lanes of a warp exchange vectors through shared memory ordered only by __syncwarp, which
// the emulation maps to a per-warp barrier — under ThreadSanitizer a missing __syncwarp is a data race even though the
// hardware would usually run the lanes in lock step.  tests/test_emulated_kernels.py: bit for bit against the CPU checker,
// and race-free.  Test infrastructure only.
#include "cuda_emu.h"

#include "../../ellalgo_rs_b200/csrc/batch_warp.cu"

namespace ellcuda {
void set_error(const std::string&) {}
namespace {
alignas(16) double smem[40000];   // the CTA's dynamic shared memory (`extern __shared__ double smem[]` in the kernel)
}
}  // namespace ellcuda

using namespace ellcuda;

constexpr int kEmuWarps = 2;      // two problems in flight per emulated CTA (the device runs 8 or 16 warps per CTA)

template <class Oracle, bool FIXED32>
static void run(const BatchArgs& a) {
    for (double& v : smem) v = NAN;               // uninitialised shared memory must never reach a result
    emu::run_cta(kEmuWarps * 32, [&](int) { batch_warp_kernel<Oracle, FIXED32, kEmuWarps, 1>(a); });
}

// same contract as ellcuda_batch_solve for host buffers; the template choices are launch_batch_warp's
extern "C" int emu_batch_warp_solve(const ell_batch_desc* d, const ell_batch_result* r) {
    if (d->kind != ELL_KIND_ELL || d->mode != ELL_MODE_OPTIM || d->n < 1 || d->n > 32) return ELLCUDA_E_ARG;
    BatchArgs a;
    std::memset(&a, 0, sizeof a);
    a.n = d->n; a.mode = d->mode; a.kind = d->kind;
    a.no_defer_trick = d->no_defer_trick; a.use_parallel_cut = d->use_parallel_cut;
    a.params_width = d->params_width;
    a.count = d->count; a.max_iters = d->max_iters; a.tolerance = d->tolerance;
    a.kappa0 = d->kappa0; a.kappa0_stride = d->kappa0_stride;
    a.xc0 = d->xc0; a.xc0_stride = d->xc0_stride;
    a.diag0 = d->diag0; a.diag0_stride = d->diag0_stride;
    a.mq0 = d->mq0; a.mq0_stride = d->mq0_stride;
    a.gamma0 = d->gamma0; a.gamma0_stride = d->gamma0_stride;
    a.params = d->params; a.params_stride = d->params_stride;
    a.table = d->table; a.table_len = d->table_len;
    for (int k = 0; k < 8; ++k) a.aux[k] = d->aux[k];
    a.xbest = r->xbest; a.has_xbest = r->has_xbest; a.num_iters = (long long*)r->num_iters;
    a.exit_status = r->exit_status; a.gamma = r->gamma; a.tsq = r->tsq; a.xc = r->xc; a.kappa = r->kappa;
    unsigned long long queue = 0;
    a.queue = &queue;
    if (d->oracle_id == ELL_ORACLE_BOWL) {
        if (a.n != 32) run<OrBowl<false>, false>(a);
        else run<OrBowl<true>, true>(a);
    } else if (d->oracle_id == ELL_ORACLE_LOWPASS) {
        if ((size_t)d->aux[0] * (d->n + 1) + kEmuWarps * kWarpScratch + 2 > sizeof smem / sizeof smem[0]) return ELLCUDA_E_ARG;
        if (a.n != 32) run<OrLowpass<false, 1>, false>(a);
        else run<OrLowpass<true, 1, false, 8>, true>(a);      // what a new_with_matrix start at n = 32 runs
    } else {
        return ELLCUDA_E_ARG;
    }
    return ELLCUDA_OK;
}

#ifdef EMU_MAIN
static int solve(const char* name, ell_batch_desc d, int count, int n) {
    std::vector<double> xbest((size_t)count * n, 0.0), gamma(count), tsq(count), xc((size_t)count * n), kappa(count);
    std::vector<uint8_t> has(count);
    std::vector<int64_t> ni(count);
    std::vector<int32_t> es(count);
    ell_batch_result r{xbest.data(), has.data(), ni.data(), es.data(), gamma.data(), tsq.data(), xc.data(), kappa.data()};
    d.count = count;
    d.n = n;
    const int rc = emu_batch_warp_solve(&d, &r);
    long long it = 0, best = 0;
    for (int i = 0; i < count; ++i) { it += ni[i]; best += has[i]; }
    std::printf("%s: rc %d, %lld iterations, %lld of %d with xbest, exit %d, gamma[0] %.17g\n", name, rc, it, best, count, es[0], gamma[0]);
    return rc;
}

int main() {
    int bad = 0;
    const double kappa100 = 100.0, one = 1.0, inf = INFINITY, kappa40 = 40.0;
    for (int n : {20, 32}) {   // bowl: generic and FIXED32 variants; diagonal start, then a non-symmetric new_with_matrix start
        const int count = 5;
        std::vector<double> xc0(n, 0.0), tgt((size_t)count * n), mq((size_t)n * n);
        for (size_t k = 0; k < tgt.size(); ++k) tgt[k] = std::sin(0.37 * (double)k);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) mq[(size_t)i * n + j] = (i == j ? 100.0 : 0.0) + 0.3 * std::sin(1.0 + 3.0 * i + 7.0 * j);
        ell_batch_desc d{};
        d.kind = ELL_KIND_ELL; d.oracle_id = ELL_ORACLE_BOWL; d.mode = ELL_MODE_OPTIM;
        d.max_iters = 120; d.tolerance = 1e-10; d.use_parallel_cut = 1;
        d.kappa0 = &kappa100; d.xc0 = xc0.data(); d.gamma0 = &inf;
        d.params = tgt.data(); d.params_stride = n; d.params_width = n;
        bad |= solve(n == 32 ? "bowl n=32 diagonal" : "bowl n=20 diagonal", d, count, n);
        d.kappa0 = &one; d.mq0 = mq.data(); d.no_defer_trick = 1;
        bad |= solve(n == 32 ? "bowl n=32 matrix start" : "bowl n=20 matrix start", d, count, n);
    }
    for (int n : {12, 32}) {   // lowpass: spectrum staged into shared memory by the whole CTA (__syncthreads), then warp-local
        const int count = 3, mdim = 15 * n;
        std::vector<double> table((size_t)mdim * n), xc0(n, 0.0), params((size_t)count * 5), gamma0(count);
        for (int i = 0; i < mdim; ++i) {
            const double w = M_PI * (double)i / (double)(mdim - 1);
            table[(size_t)i * n] = 1.0;
            for (int j = 1; j < n; ++j) table[(size_t)i * n + j] = 2.0 * std::cos(w * (double)j);
        }
        for (int i = 0; i < count; ++i) {
            const double wpass = 0.12 + 0.01 * i, wstop = wpass + 0.08, dp = 0.025, ds = 0.125;
            const double d1 = 20.0 * std::log10(1.0 + dp), lp = std::pow(10.0, -d1 / 20.0), up = std::pow(10.0, d1 / 20.0);
            double* p = params.data() + 5 * i;
            p[0] = wpass; p[1] = wstop; p[2] = lp * lp; p[3] = up * up; p[4] = ds * ds;
            gamma0[i] = p[4];
        }
        ell_batch_desc d{};
        d.kind = ELL_KIND_ELL; d.oracle_id = ELL_ORACLE_LOWPASS; d.mode = ELL_MODE_OPTIM;
        d.max_iters = 400; d.tolerance = 1e-14; d.use_parallel_cut = 1;
        d.kappa0 = &kappa40; d.xc0 = xc0.data(); d.gamma0 = gamma0.data(); d.gamma0_stride = 1;
        d.params = params.data(); d.params_stride = 5; d.params_width = 5;
        d.table = table.data(); d.table_len = (int64_t)table.size(); d.aux[0] = mdim;
        bad |= solve(n == 32 ? "lowpass n=32" : "lowpass n=12", d, count, n);
    }
    return bad;
}
#endif
