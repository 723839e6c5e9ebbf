// emu_ellcta.cpp — the CTA-per-problem kernel's DEVICE CODE (ellalgo_rs_b200/csrc/ellcta_device.cuh) compiled for the host:
// one OS thread per CUDA thread of a CTA, pthread barriers for __syncthreads, and per-warp barriers behind __ballot_sync /
// __shfl_xor_sync.  Built twice by tests/test_ellcta_emulated.py:
//   * plain: the kernel's logic (cursor handling, first-violator selection, packed-triangle indexing, balanced rank-1) must
//     reproduce the CPU checker bit for bit on a CPU-only box;
//   * -fsanitize=thread: OS threads really run asynchronously, so a shared-memory access that the kernel does not order
//     with a barrier is reported by ThreadSanitizer as a data race (SURVEY.md §5 "race detection"; compute-sanitizer's
//     racecheck is closed on the GPU pool).
// Test infrastructure only: nothing in the product links this.
#include "cuda_emu.h"

#include "../../ellalgo_rs_b200/csrc/ellcta_device.cuh"

namespace ellcuda {
void set_error(const std::string&) {}
}  // namespace ellcuda

using namespace ellcuda;
using namespace ellcuda::ellcta;

template <class Oracle>
static void run_grid(const BatchArgs& a, int ctas) {
    const int n = a.n, nt = ((n + 31) / 32) * 32;
    const size_t doubles = (((size_t)n * (n + 1) / 2 + 1) & ~(size_t)1) + 4 * (size_t)nt;
    std::vector<std::thread> grid;
    for (int b = 0; b < ctas; ++b)
        grid.emplace_back([&a, nt, doubles] {
            std::vector<double> smem(doubles, NAN);    // uninitialised shared memory must never reach a result
            CtaShared sh;
            std::memset(&sh, 0xff, sizeof sh);
            emu::run_cta(nt, [&](int t) { ellcta_body<Oracle>(a, smem.data(), sh, t, nt); });
        });
    for (auto& g : grid) g.join();
}

// same contract as ellcuda_batch_solve for host buffers, KIND_ELL + MODE_OPTIM + BOWL / LOWPASS (what batch_ellcta.cu launches)
extern "C" int emu_ellcta_solve(const ell_batch_desc* d, const ell_batch_result* r, int ctas) {
    if (d->kind != ELL_KIND_ELL || d->mode != ELL_MODE_OPTIM || d->n < 1 || d->n > kCtaNMax) return ELLCUDA_E_ARG;
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
    if (ctas < 1) ctas = 1;
    if (d->oracle_id == ELL_ORACLE_BOWL) {
        run_grid<OrBowlCta>(a, ctas);
    } else if (d->oracle_id == ELL_ORACLE_LOWPASS) {
        const int rows = d->aux[0], cols = d->n;      // launch_transpose_table
        std::vector<double> tt((size_t)rows * cols);
        for (int i = 0; i < rows; ++i)
            for (int j = 0; j < cols; ++j) tt[(size_t)j * rows + i] = d->table[(size_t)i * cols + j];
        a.table_t = tt.data();
        run_grid<OrLowpassCta>(a, ctas);
    } else {
        return ELLCUDA_E_ARG;
    }
    return ELLCUDA_OK;
}

#ifdef EMU_MAIN
// Stand-alone driver for the ThreadSanitizer build: a few small batches that walk every path of the kernel (diagonal and
// non-symmetric matrix starts, no_defer_trick, pass- / stop- / transition-band cuts, the stop-band maximum and the shrink),
// two CTAs at a time so that the work queue is contended.  Prints what happened; ThreadSanitizer makes the exit code non-zero
// if it saw a race.
static int run(const char* name, ell_batch_desc d, int count, int n) {
    std::vector<double> xbest((size_t)count * n, 0.0), gamma(count), tsq(count), xc((size_t)count * n), kappa(count);
    std::vector<uint8_t> has(count);
    std::vector<int64_t> ni(count);
    std::vector<int32_t> es(count);
    ell_batch_result r{xbest.data(), has.data(), ni.data(), es.data(), gamma.data(), tsq.data(), xc.data(), kappa.data()};
    d.count = count;
    d.n = n;
    const int rc = emu_ellcta_solve(&d, &r, 2);
    long long it = 0, best = 0;
    for (int i = 0; i < count; ++i) { it += ni[i]; best += has[i]; }
    std::printf("%s: rc %d, %lld iterations, %lld of %d with xbest, exit %d, gamma[0] %.17g\n", name, rc, it, best, count, es[0], gamma[0]);
    return rc;
}

int main() {
    int bad = 0;
    const double kappa100 = 100.0, one = 1.0, inf = INFINITY;
    {   // bowl, n = 40 (2 warps), diagonal start
        const int n = 40, count = 5;
        std::vector<double> xc0(n, 0.0), tgt((size_t)count * n);
        for (size_t k = 0; k < tgt.size(); ++k) tgt[k] = std::sin(0.37 * (double)k);
        ell_batch_desc d{};
        d.kind = ELL_KIND_ELL; d.oracle_id = ELL_ORACLE_BOWL; d.mode = ELL_MODE_OPTIM;
        d.max_iters = 150; d.tolerance = 1e-10; d.use_parallel_cut = 1;
        d.kappa0 = &kappa100; d.xc0 = xc0.data(); d.gamma0 = &inf;
        d.params = tgt.data(); d.params_stride = n; d.params_width = n;
        bad |= run("bowl n=40 diagonal", d, count, n);
        d.no_defer_trick = 1;
        bad |= run("bowl n=40 diagonal no_defer", d, count, n);
    }
    {   // bowl, n = 70 (3 warps, last one partly idle), non-symmetric new_with_matrix start
        const int n = 70, count = 3;
        std::vector<double> xc0(n, 0.0), tgt((size_t)count * n), mq((size_t)n * n);
        for (size_t k = 0; k < tgt.size(); ++k) tgt[k] = std::cos(0.11 * (double)k);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) mq[(size_t)i * n + j] = (i == j ? 100.0 : 0.0) + 0.3 * std::sin(1.0 + 3.0 * i + 7.0 * j);
        ell_batch_desc d{};
        d.kind = ELL_KIND_ELL; d.oracle_id = ELL_ORACLE_BOWL; d.mode = ELL_MODE_OPTIM;
        d.max_iters = 80; d.tolerance = 1e-10; d.use_parallel_cut = 1;
        d.kappa0 = &one; d.xc0 = xc0.data(); d.gamma0 = &inf; d.mq0 = mq.data();
        d.params = tgt.data(); d.params_stride = n; d.params_width = n;
        bad |= run("bowl n=70 non-symmetric start", d, count, n);
    }
    {   // lowpass, n = 40, spectrum as LowpassOracle::new builds it (src/oracles/lowpass_oracle.rs:24-33), sane bounds
        const int n = 40, count = 3, mdim = 15 * n;
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
        const double kappa40 = 40.0;
        ell_batch_desc d{};
        d.kind = ELL_KIND_ELL; d.oracle_id = ELL_ORACLE_LOWPASS; d.mode = ELL_MODE_OPTIM;
        d.max_iters = 500; d.tolerance = 1e-14; d.use_parallel_cut = 1;
        d.kappa0 = &kappa40; d.xc0 = xc0.data(); d.gamma0 = gamma0.data(); d.gamma0_stride = 1;
        d.params = params.data(); d.params_stride = 5; d.params_width = 5;
        d.table = table.data(); d.table_len = (int64_t)table.size(); d.aux[0] = mdim;
        bad |= run("lowpass n=40", d, count, n);
    }
    return bad;
}
#endif
