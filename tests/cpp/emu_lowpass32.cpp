// emu_lowpass32.cpp — the HEADLINE kernel (ellalgo_rs_b200/csrc/batch_lowpass.cu: LowpassOracle with 32 taps under
// cutting_plane_optim, one warp per problem, M in tensor memory, residue-mapped rounds over the spectrum in shared memory)
// compiled for the host with cuda_emu.h + tmem_emu.h.  The .cu file is included as it is (host launchers behind
// #ifdef __CUDACC__).  tests/test_emulated_kernels.py: bit for bit against the CPU checker on a CPU-only box (the rounds that
// never cross a band's wrap-around point, the rotated ballot, the cursors), and race-free under ThreadSanitizer — shared
// memory between lanes, tensor-memory cells between warps.  Test infrastructure only.
#include "cuda_emu.h"
#include "tmem_emu.h"

#include "../../ellalgo_rs_b200/csrc/batch_lowpass.cu"

namespace ellcuda {
void set_error(const std::string&) {}
namespace {
alignas(16) double smem[40000];   // the CTA's dynamic shared memory (`extern __shared__ double smem[]` in the kernel)
}
}  // namespace ellcuda

using namespace ellcuda;

constexpr int kEmuWarps = 6;      // warps 4 and 5 share lane quarters 0 and 1 with warps 0 and 1, in the next column block

// same contract as ellcuda_batch_solve for host buffers; variant as ELLCUDA_LOWPASS_VARIANT (0 = what bench.py times)
extern "C" int emu_lowpass32_solve(const ell_batch_desc* d, const ell_batch_result* r, int variant) {
    if (d->kind != ELL_KIND_ELL || d->mode != ELL_MODE_OPTIM || d->oracle_id != ELL_ORACLE_LOWPASS || d->n != kN ||
        d->aux[0] != 15 * kN || d->mq0)
        return ELLCUDA_E_ARG;
    BatchArgs a;
    std::memset(&a, 0, sizeof a);
    a.n = d->n; a.mode = d->mode; a.kind = d->kind;
    a.no_defer_trick = d->no_defer_trick; a.use_parallel_cut = d->use_parallel_cut;
    a.params_width = d->params_width;
    a.count = d->count; a.max_iters = d->max_iters; a.tolerance = d->tolerance;
    a.kappa0 = d->kappa0; a.kappa0_stride = d->kappa0_stride;
    a.xc0 = d->xc0; a.xc0_stride = d->xc0_stride;
    a.diag0 = d->diag0; a.diag0_stride = d->diag0_stride;
    a.gamma0 = d->gamma0; a.gamma0_stride = d->gamma0_stride;
    a.params = d->params; a.params_stride = d->params_stride;
    a.table = d->table; a.table_len = d->table_len;
    for (int k = 0; k < 8; ++k) a.aux[k] = d->aux[k];
    a.xbest = r->xbest; a.has_xbest = r->has_xbest; a.num_iters = (long long*)r->num_iters;
    a.exit_status = r->exit_status; a.gamma = r->gamma; a.tsq = r->tsq; a.xc = r->xc; a.kappa = r->kappa;
    unsigned long long queue = 0;
    a.queue = &queue;
    for (double& v : smem) v = NAN;               // uninitialised shared / tensor memory must never reach a result
    std::memset(emu_tmem, 0xff, sizeof emu_tmem);
    if (variant == 1) emu::run_cta(kEmuWarps * 32, [&](int) { lowpass32_tmem_kernel<8, false, kEmuWarps>(a); });
    else if (variant == 2) emu::run_cta(kEmuWarps * 32, [&](int) { lowpass32_tmem_kernel<8, true, kEmuWarps>(a); });
    else emu::run_cta(kEmuWarps * 32, [&](int) { lowpass32_tmem_kernel<0, false, kEmuWarps>(a); });
    return ELLCUDA_OK;
}

#ifdef EMU_MAIN
int main() {
    const int n = kN, count = 9, mdim = 15 * n;     // more problems than warps: the work queue hands out a second round
    std::vector<double> table((size_t)mdim * n), xc0(n, 0.0), params((size_t)count * 5), gamma0(count);
    for (int i = 0; i < mdim; ++i) {                // LowpassOracle::new's spectrum (src/oracles/lowpass_oracle.rs:24-33)
        const double w = M_PI * (double)i / (double)(mdim - 1);
        table[(size_t)i * n] = 1.0;
        for (int j = 1; j < n; ++j) table[(size_t)i * n + j] = 2.0 * std::cos(w * (double)j);
    }
    for (int i = 0; i < count; ++i) {
        const double wpass = 0.10 + 0.005 * i, wstop = wpass + 0.08, dp = 0.025, ds = 0.125;
        const double d1 = 20.0 * std::log10(1.0 + dp), lp = std::pow(10.0, -d1 / 20.0), up = std::pow(10.0, d1 / 20.0);
        double* p = params.data() + 5 * i;
        p[0] = wpass; p[1] = wstop; p[2] = lp * lp; p[3] = up * up; p[4] = ds * ds;
        gamma0[i] = p[4];
    }
    const double kappa40 = 40.0;
    ell_batch_desc d{};
    d.kind = ELL_KIND_ELL; d.oracle_id = ELL_ORACLE_LOWPASS; d.mode = ELL_MODE_OPTIM;
    d.n = n; d.count = count; d.max_iters = 500; d.tolerance = 1e-14; d.use_parallel_cut = 1;
    d.kappa0 = &kappa40; d.xc0 = xc0.data(); d.gamma0 = gamma0.data(); d.gamma0_stride = 1;
    d.params = params.data(); d.params_stride = 5; d.params_width = 5;
    d.table = table.data(); d.table_len = (int64_t)table.size(); d.aux[0] = mdim;
    int bad = 0;
    for (int variant : {0, 1}) {
        std::vector<double> xbest((size_t)count * n, 0.0), gamma(count), tsq(count), xc((size_t)count * n), kappa(count);
        std::vector<uint8_t> has(count);
        std::vector<int64_t> ni(count);
        std::vector<int32_t> es(count);
        ell_batch_result r{xbest.data(), has.data(), ni.data(), es.data(), gamma.data(), tsq.data(), xc.data(), kappa.data()};
        bad |= emu_lowpass32_solve(&d, &r, variant);
        long long it = 0, best = 0;
        for (int i = 0; i < count; ++i) { it += ni[i]; best += has[i]; }
        std::printf("lowpass32 variant %d: %lld iterations, %lld of %d with xbest, exit %d, gamma[0] %.17g\n", variant, it, best, count, es[0], gamma[0]);
    }
    return bad;
}
#endif
