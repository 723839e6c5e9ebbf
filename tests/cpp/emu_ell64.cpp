// emu_ell64.cpp — the batched `Ell` kernel with M in tensor memory (ellalgo_rs_b200/csrc/batch_ell64.cu: n <= 64, one warp per
// problem, M streamed through registers in 16-double chunks; oracles: bowl, LowpassOracle with the spectrum transposed in
// global memory, LMIOracle / LMIOldOracle / LMI0Oracle on the `Ell` space) compiled for the host with cuda_emu.h +
// tmem_emu.h, 6 warps per CTA instead of 8.  The .cu file is included as it is (host launchers behind #ifdef __CUDACC__).
// tests/test_emulated_kernels.py: bit for bit against the CPU checker, race-free under ThreadSanitizer.  Test infrastructure only.
#define ELLCUDA_ELL64_WARPS 6
#include "cuda_emu.h"
#include "tmem_emu.h"

#include "../../ellalgo_rs_b200/csrc/batch_ell64.cu"

namespace ellcuda {
void set_error(const std::string&) {}
namespace {
alignas(16) double smem[40000];   // the CTA's dynamic shared memory (`extern __shared__ double smem[]` in the kernel)
}
}  // namespace ellcuda

using namespace ellcuda;

// same contract as ellcuda_batch_solve for host buffers, KIND_ELL; the shared-memory layout figures are launch_batch_ell64's
extern "C" int emu_ell64_solve(const ell_batch_desc* d, const ell_batch_result* r) {
    const int n = d->n;
    if (d->kind != ELL_KIND_ELL || n < 1 || n > kNMax) return ELLCUDA_E_ARG;
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
    a.gamma0 = d->mode == ELL_MODE_FEAS ? nullptr : d->gamma0; a.gamma0_stride = d->gamma0_stride;
    a.params = d->params; a.params_stride = d->params_stride;
    a.table = d->table; a.table_len = d->table_len;
    for (int k = 0; k < 8; ++k) a.aux[k] = d->aux[k];
    a.xbest = r->xbest; a.has_xbest = r->has_xbest; a.num_iters = (long long*)r->num_iters;
    a.exit_status = r->exit_status; a.gamma = r->gamma; a.tsq = r->tsq; a.xc = r->xc; a.kappa = r->kappa;
    unsigned long long queue = 0;
    a.queue = &queue;
    for (double& v : smem) v = NAN;               // uninitialised shared / tensor memory must never reach a result
    std::memset(emu_tmem, 0xff, sizeof emu_tmem);
    std::vector<double> tt;
    const int nt = kWarps * 32;
    if (d->oracle_id == ELL_ORACLE_BOWL && d->mode == ELL_MODE_OPTIM) {
        emu::run_cta(nt, [&](int) { batch_ell64_kernel<OrBowl64>(a, 0, kVec, 0); });
    } else if (d->oracle_id == ELL_ORACLE_LOWPASS && d->mode == ELL_MODE_OPTIM) {
        const int rows = d->aux[0];
        tt.resize((size_t)rows * n);                // launch_transpose_table
        for (int i = 0; i < rows; ++i)
            for (int j = 0; j < n; ++j) tt[(size_t)j * rows + i] = d->table[(size_t)i * n + j];
        a.table_t = tt.data();
        emu::run_cta(nt, [&](int) { batch_ell64_kernel<OrLowpass64>(a, 0, kVec, 0); });
    } else if ((d->oracle_id == ELL_ORACLE_LMI || d->oracle_id == ELL_ORACLE_LMI_OLD || d->oracle_id == ELL_ORACLE_LMI0) &&
               d->mode == ELL_MODE_FEAS) {
        const int lmi0 = d->oracle_id == ELL_ORACLE_LMI0, m = d->aux[0];
        if (m < 1 || m > 32) return ELLCUDA_E_ARG;
        const int tri = m * (m + 1) / 2, fstride = tri | 1;
        int per_warp = kVec + ((tri + 1) & ~1) + m * m + ((m + 1) & ~1);
        per_warp = (per_warp + 1) & ~1;
        const size_t fixed = ((size_t)n * fstride + tri + 1) & ~(size_t)1;
        if (fixed + (size_t)kWarps * per_warp > sizeof smem / sizeof smem[0]) return ELLCUDA_E_ARG;
        emu::run_cta(nt, [&](int) { batch_ell64_kernel<OrLmi64>(a, fstride, per_warp, lmi0); });
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
    const int rc = emu_ell64_solve(&d, &r);
    long long it = 0, best = 0;
    for (int i = 0; i < count; ++i) { it += ni[i]; best += has[i]; }
    std::printf("%s: rc %d, %lld iterations, %lld of %d with xbest, exit[0] %d\n", name, rc, it, best, count, es[0]);
    return rc;
}

int main() {
    int bad = 0;
    const double kappa100 = 100.0, one = 1.0, inf = INFINITY, kappa40 = 40.0;
    {   // bowl n = 48: diagonal start, then a non-symmetric new_with_matrix start with no_defer_trick
        const int n = 48, count = 8;
        std::vector<double> xc0(n, 0.0), tgt((size_t)count * n), mq((size_t)n * n);
        for (size_t k = 0; k < tgt.size(); ++k) tgt[k] = std::sin(0.37 * (double)k);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) mq[(size_t)i * n + j] = (i == j ? 100.0 : 0.0) + 0.3 * std::sin(1.0 + 3.0 * i + 7.0 * j);
        ell_batch_desc d{};
        d.kind = ELL_KIND_ELL; d.oracle_id = ELL_ORACLE_BOWL; d.mode = ELL_MODE_OPTIM;
        d.max_iters = 60; d.tolerance = 1e-10; d.use_parallel_cut = 1;
        d.kappa0 = &kappa100; d.xc0 = xc0.data(); d.gamma0 = &inf;
        d.params = tgt.data(); d.params_stride = n; d.params_width = n;
        bad |= solve("bowl n=48 diagonal", d, count, n);
        d.kappa0 = &one; d.mq0 = mq.data(); d.no_defer_trick = 1;
        bad |= solve("bowl n=48 matrix start", d, count, n);
    }
    {   // lowpass n = 36
        const int n = 36, count = 3, mdim = 15 * n;
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
        d.max_iters = 500; d.tolerance = 1e-14; d.use_parallel_cut = 1;
        d.kappa0 = &kappa40; d.xc0 = xc0.data(); d.gamma0 = gamma0.data(); d.gamma0_stride = 1;
        d.params = params.data(); d.params_stride = 5; d.params_width = 5;
        d.table = table.data(); d.table_len = (int64_t)table.size(); d.aux[0] = mdim;
        bad |= solve("lowpass n=36", d, count, n);
    }
    {   // LMI0 on Ell, n = 40, m = 9
        const int n = 40, m = 9, count = 7;
        std::vector<double> table((size_t)n * m * m), xc0((size_t)count * n);
        for (int k = 0; k < n; ++k)
            for (int i = 0; i < m; ++i)
                for (int j = 0; j <= i; ++j) {
                    double v = std::sin(1.0 + 0.7 * k + 1.3 * i + 2.9 * j);
                    if (k == 0) v = (i == j ? 1.0 : 0.0) + 0.25 * v;
                    table[(size_t)k * m * m + i * m + j] = table[(size_t)k * m * m + j * m + i] = v;
                }
        for (size_t k = 0; k < xc0.size(); ++k) xc0[k] = 0.2 * std::cos(0.9 * (double)k);
        ell_batch_desc d{};
        d.kind = ELL_KIND_ELL; d.oracle_id = ELL_ORACLE_LMI0; d.mode = ELL_MODE_FEAS;
        d.max_iters = 200; d.tolerance = 1e-20; d.use_parallel_cut = 1;
        d.kappa0 = &kappa100; d.xc0 = xc0.data(); d.xc0_stride = n;
        d.table = table.data(); d.table_len = (int64_t)table.size(); d.aux[0] = m;
        bad |= solve("LMI0 on Ell n=40 m=9", d, count, n);
    }
    return bad;
}
#endif
