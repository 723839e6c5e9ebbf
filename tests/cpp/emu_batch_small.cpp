// emu_batch_small.cpp — the n = 2 kernels of ellalgo_rs_b200/csrc/batch_small.cu (one thread per problem: ProfitOracle /
// ProfitRbOracle / ProfitOracleQ with the glibc-exact exp and ln of glibc_math.cuh, quasicvx, example1 / 1_rr / 3 / 4, the
// cutting_plane_tests fixtures; `Ell` and `EllStable`; optim, feas, optim_q and bsearch drivers) compiled for the host with
// cuda_emu.h.  The .cu file is included as it is — its own dispatch table (launch_batch_small) picks the instantiation and
// only the launch is replaced by an emulated CTA of 128 threads.  tests/test_emulated_kernels.py: every iteration count the
// reference's tests pin (25, 25, 82, 83, 90, 29, 35, (true, 34), (None, 2) …) comes out of the kernels' own code on a
// CPU-only box.  Test infrastructure only.
#include "cuda_emu.h"

#include "../../ellalgo_rs_b200/csrc/common.cuh"
namespace ellcuda {
namespace {
template <class Oracle, int MODE>
int launch_kind(const BatchArgs& a, cudaStream_t s, int sm_count);   // defined below, after the kernels
}
}  // namespace ellcuda

#include "../../ellalgo_rs_b200/csrc/batch_small.cu"

namespace ellcuda {
static std::string emu_last_error;
void set_error(const std::string& m) { emu_last_error = m; }
namespace {
template <class Oracle, int MODE>
int launch_kind(const BatchArgs& a, cudaStream_t, int) {
    constexpr int N = Oracle::kN;
    if (a.kind == ELL_KIND_ELLSTABLE) emu::run_cta(128, [&](int) { batch_small_kernel<N, EllStableReg<N>, Oracle, MODE>(a); });
    else emu::run_cta(128, [&](int) { batch_small_kernel<N, EllReg<N>, Oracle, MODE>(a); });
    return ELLCUDA_OK;
}
}  // namespace
}  // namespace ellcuda

using namespace ellcuda;

// same contract as ellcuda_batch_solve for host buffers and the oracles of the n = 2 kernel
extern "C" int emu_batch_small_solve(const ell_batch_desc* d, const ell_batch_result* r) {
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
    long long launches = 0;
    return launch_batch_small(a, d->oracle_id, nullptr, 1, &launches);
}
