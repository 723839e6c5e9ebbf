// emu_lmi.cpp — the batched EllStable + LMI kernel of ellalgo_rs_b200/csrc/batch_stable.cu (configs[3]: LMIOracle / LMIOldOracle /
// LMI0Oracle with ldlt_mgr under cutting_plane_feas, one warp per problem, DIAG and GENERAL factor variants; lmi_device.cuh
// holds the oracle: two LDL^T rows per step in half-warps, lane-parallel witness products) compiled for the host with
// cuda_emu.h.  The .cu file is included as it is (its host launchers sit behind #ifdef __CUDACC__).  The lanes of a warp
// exchange the factor rows, the witness and the EllStable vectors through shared memory ordered only by __syncwarp —
// 24 of them per iteration — which the emulation maps to a per-warp barrier: under ThreadSanitizer a missing one is a data
// race.  tests/test_emulated_kernels.py: bit for bit against the CPU checker, and race-free.  Test infrastructure only.
#include "cuda_emu.h"

#include "../../ellalgo_rs_b200/csrc/batch_stable.cu"

namespace ellcuda {
void set_error(const std::string&) {}
namespace {
alignas(16) double smem[40000];   // the CTA's dynamic shared memory (`extern __shared__ double smem[]` in the kernel)
}
}  // namespace ellcuda

using namespace ellcuda;

constexpr int kEmuWarps = 2;      // two problems in flight per emulated CTA (the device runs up to 28 warps per CTA)

// same contract as ellcuda_batch_solve for host buffers, KIND_ELLSTABLE + MODE_FEAS + LMI / LMI_OLD / LMI0; the shared-memory
// layout figures are launch_batch_stable's
extern "C" int emu_lmi_solve(const ell_batch_desc* d, const ell_batch_result* r) {
    const int n = d->n, m = d->aux[0];
    const int lmi0 = d->oracle_id == ELL_ORACLE_LMI0;
    if (d->kind != ELL_KIND_ELLSTABLE || d->mode != ELL_MODE_FEAS || n < 1 || n > kNMax || m < 1 || m > 32) return ELLCUDA_E_ARG;
    if (d->oracle_id != ELL_ORACLE_LMI && d->oracle_id != ELL_ORACLE_LMI_OLD && !lmi0) return ELLCUDA_E_ARG;
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
    a.params = d->params; a.params_stride = d->params_stride;
    a.table = d->table; a.table_len = d->table_len;
    for (int k = 0; k < 8; ++k) a.aux[k] = d->aux[k];
    a.xbest = r->xbest; a.has_xbest = r->has_xbest; a.num_iters = (long long*)r->num_iters;
    a.exit_status = r->exit_status; a.gamma = r->gamma; a.tsq = r->tsq; a.xc = r->xc; a.kappa = r->kappa;
    unsigned long long queue = 0;
    a.queue = &queue;
    const int tri = m * (m + 1) / 2, fstride = tri | 1;
    const bool diag = a.mq0 == nullptr;
    int per_warp = 4 * kNMax + ((tri + 1) & ~1) + m * m + ((m + 1) & ~1);
    if (!diag) per_warp += (n * (n + 1) / 2 + 1) & ~1;
    per_warp = (per_warp + 1) & ~1;
    const size_t fixed = ((size_t)n * fstride + tri + 1) & ~(size_t)1;
    if (fixed + (size_t)kEmuWarps * per_warp > sizeof smem / sizeof smem[0]) return ELLCUDA_E_ARG;
    for (double& v : smem) v = NAN;               // uninitialised shared memory must never reach a result
    if (diag) emu::run_cta(kEmuWarps * 32, [&](int) { batch_stable_kernel<true, 512>(a, kEmuWarps, fstride, per_warp, lmi0); });
    else emu::run_cta(kEmuWarps * 32, [&](int) { batch_stable_kernel<false, 512>(a, kEmuWarps, fstride, per_warp, lmi0); });
    return ELLCUDA_OK;
}

// one update_core of the single EllStable handle with n <= 64 (stable_single_kernel: one warp, the packed factor staged in
// static shared memory); arguments as emu_stable_large_update
extern "C" int emu_stable_single_update(double* qu, double* xc, double* scal, int* status, const double* g, int n, int use_parallel,
                                        int which, double beta0, int has_beta1, double beta1) {
    if (n < 1 || n > kNMax) return ELLCUDA_E_ARG;
    const CalcConsts cc = CalcConsts::make(n, use_parallel);
    const Cut cut = has_beta1 > 0 ? parallel_cut(beta0, beta1) : single_cut(beta0);
    emu::run_cta(32, [&](int) { stable_single_kernel(qu, xc, scal, status, g, n, cc, which, cut); });
    return ELLCUDA_OK;
}

#ifdef EMU_MAIN
static int solve(const char* name, ell_batch_desc d, int count, int n) {
    std::vector<double> xbest((size_t)count * n, 0.0), tsq(count), xc((size_t)count * n), kappa(count);
    std::vector<uint8_t> has(count);
    std::vector<int64_t> ni(count);
    std::vector<int32_t> es(count);
    ell_batch_result r{xbest.data(), has.data(), ni.data(), es.data(), nullptr, tsq.data(), xc.data(), kappa.data()};
    d.count = count;
    d.n = n;
    const int rc = emu_lmi_solve(&d, &r);
    long long it = 0, feas = 0;
    for (int i = 0; i < count; ++i) { it += ni[i]; feas += es[i] == 6; }
    std::printf("%s: rc %d, %lld iterations, %lld of %d feasible, exit[0] %d, kappa[0] %.17g\n", name, rc, it, feas, count, es[0], kappa[0]);
    return rc;
}

int main() {
    int bad = 0;
    const int n = 24, m = 9, count = 4;
    std::vector<double> table((size_t)(n + 1) * m * m), xc0((size_t)count * n), mq((size_t)n * n, 0.0);
    for (int k = 0; k <= n; ++k)                  // symmetric F_k, then B = I + small symmetric
        for (int i = 0; i < m; ++i)
            for (int j = 0; j <= i; ++j) {
                double v = std::sin(1.0 + 0.7 * k + 1.3 * i + 2.9 * j);
                if (k == n) v = (i == j ? 1.0 : 0.0) + 0.0125 * v;
                table[(size_t)k * m * m + i * m + j] = table[(size_t)k * m * m + j * m + i] = v;
            }
    for (size_t k = 0; k < xc0.size(); ++k) xc0[k] = 0.2 * std::cos(0.9 * (double)k);
    for (int i = 0; i < n; ++i)
        for (int j = i; j < n; ++j) mq[(size_t)i * n + j] = i == j ? 1.0 : 1e-3 * std::sin(2.0 + i + 5.0 * j);
    const double kappa100 = 100.0;
    ell_batch_desc d{};
    d.kind = ELL_KIND_ELLSTABLE; d.mode = ELL_MODE_FEAS;
    d.max_iters = 150; d.tolerance = 1e-20; d.use_parallel_cut = 1;
    d.kappa0 = &kappa100; d.xc0 = xc0.data(); d.xc0_stride = n;
    d.table = table.data(); d.table_len = (int64_t)table.size(); d.aux[0] = m;
    d.oracle_id = ELL_ORACLE_LMI;
    bad |= solve("LMI diagonal start", d, count, n);
    d.oracle_id = ELL_ORACLE_LMI_OLD;
    bad |= solve("LMIOld diagonal start", d, count, n);
    d.oracle_id = ELL_ORACLE_LMI;
    d.mq0 = mq.data();
    bad |= solve("LMI general start", d, count, n);
    d.mq0 = nullptr;
    d.oracle_id = ELL_ORACLE_LMI0;
    for (int i = 0; i < m; ++i)                   // F_0 = I + 0.25 F_0 keeps the cone non-trivial
        for (int j = 0; j < m; ++j) table[(size_t)i * m + j] = (i == j ? 1.0 : 0.0) + 0.25 * table[(size_t)i * m + j];
    d.table_len = (int64_t)n * m * m;
    bad |= solve("LMI0 diagonal start", d, count, n);
    {   // the single handle (n <= 64): three updates of an n = 40 factor, one of them failing
        const int ns = 40;
        std::vector<double> qu((size_t)ns * (ns + 1) / 2), xs(ns), g(ns), scal{2.5, 0.0};
        size_t e = 0;
        for (int r = 0; r < ns; ++r)
            for (int c = r; c < ns; ++c) qu[e++] = c == r ? 1.0 + 0.01 * r : 1e-3 * std::sin(1.0 + 3.0 * r + 7.0 * c);
        for (int i = 0; i < ns; ++i) xs[i] = std::cos(0.3 * i);
        int status = -1, step = 0;
        const struct { int which; double b0; int has; double b1; } cuts[] = {{0, 0.01, 1, 0.10}, {0, 1e6, -1, 0.0}, {1, 0.0, -1, 0.0}};
        for (const auto& c : cuts) {
            for (int i = 0; i < ns; ++i) g[i] = std::sin(0.7 * i + step) / std::sqrt((double)ns);
            bad |= emu_stable_single_update(qu.data(), xs.data(), scal.data(), &status, g.data(), ns, 1, c.which, c.b0, c.has, c.b1);
            std::printf("single handle step %d: status %d, kappa %.17g\n", step++, status, scal[0]);
        }
    }
    return bad;
}
#endif
