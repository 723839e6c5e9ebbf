// emu_stream.cpp — the LARGE-n path of ellalgo_rs_b200/csrc/stream_update.cu (sweep_kernel: deferred rank-1 + column sums of M
// through a TMA / mbarrier ring, segment combine across CTAs, plain or flagged-line output; finish_kernel: omega in the
// reference's order or as a fixed tree, EllCalc, xc, the vectors of the next deferred rank-1) compiled for the host with
// cuda_emu.h + stream_ptx_emu.h.  The .cu file is included as it is (its host side sits behind #ifdef __CUDACC__); what this
// file adds is the handle's bookkeeping — the buffers and launch shapes of alloc_ctx / enqueue_begin / ellcuda_update_finish,
// restated for host memory — and a grid loop that runs the CTAs one after the other.
// tests/test_emulated_kernels.py: Ell::update_core on the emulated kernels equals the CPU checker bit for bit (reduce_mode 0)
// for column-block widths 64 / 32 / 16, one or two warps per CTA, row segments, no_defer_trick, dense symmetric starts, failing
// cuts, with and without the flagged-line exchange; and the hand-offs (mbarrier phases, the CTA-wide staging of the ordered
// sum) are race-free under ThreadSanitizer.  Test infrastructure only.
// finish_kernel keeps the device's 1024 threads (and with them the device's summation tree in reduce_mode > 0) in the plain
// build; the ThreadSanitizer build passes -DELLCUDA_FIN_THREADS=128 (the sanitizer follows <= 256 running threads exactly).
#include "cuda_emu.h"
#include "stream_ptx_emu.h"

#include "../../ellalgo_rs_b200/csrc/stream_update.cu"

namespace ellcuda {
void set_error(const std::string&) {}
void keep_mempool(int) {}
namespace {
alignas(128) unsigned char smem_raw[128 * 1024];   // the sweep CTA's dynamic shared memory (`extern __shared__ ... smem_raw[]`)
}
}  // namespace ellcuda

using namespace ellcuda;

namespace {

struct EmuEll {   // the fields of ellcuda_ctx the kernels see (full storage, one slab = the whole matrix)
    long long n, w, c0, ld, npad, ld_part, ncb;
    int cols, tile_rows, stages, nseg, reduce_mode, no_defer, use_parallel, nranks, warps, order;
    std::vector<double> S, xc, gpad, gzero, partial, xblk;
    std::vector<double2> pend;
    std::vector<unsigned> counters;
    Scal sc;
    CUtensorMap map;
    int rank = 0;
    std::vector<EmuEll*> peers;   // row-sharded group: every rank's handle (this one included), else empty
};

template <int CO, int R, int ST, int W>
void sweep_grid(EmuEll* h, const SweepArgs& a) {
    static_assert(sweep_smem(CO, R, ST) <= sizeof smem_raw, "");
    for (long long k = 0; k < h->ncb * h->nseg; ++k) {
        // the order in which CTAs run decides which row segment arrives last and sums the block: alternate it
        const long long kk = (h->order & 1) ? h->ncb * h->nseg - 1 - k : k;
        const unsigned cb = (unsigned)(kk % h->ncb), seg = (unsigned)(kk / h->ncb);
        std::memset(smem_raw, 0xff, sizeof smem_raw);
        emu::run_cta(32 * W, [&](int) { sweep_kernel<CO, R, ST, W>(h->map, a); }, EmuDim3{cb, seg, 0}, EmuDim3{(unsigned)h->ncb, (unsigned)h->nseg, 1});
    }
    h->order += 1;
}

void run_sweep(EmuEll* h, const double* g, double* plain_out) {   // enqueue_begin; g == nullptr: flush
    SweepArgs a;
    a.g = g ? g : h->gzero.data(); a.pend = h->pend.data(); a.sc = &h->sc; a.partial = h->partial.data();
    a.cnt_block = h->counters.data(); a.own_blk = h->xblk.data(); a.out = plain_out;
    a.ld_part = h->ld_part; a.n = h->n; a.w = h->w; a.c_glob0 = h->c0; a.nseg = h->nseg;
    a.exchange = g ? 1 : 0;
    a.pt.nranks = h->nranks; a.pt.rank = h->rank;
    for (int r = 0; r < kMaxRanks; ++r)
        a.pt.blk[r] = r < h->nranks ? (h->peers.empty() ? h->xblk.data() : h->peers[r]->xblk.data()) : nullptr;
    if (h->cols == 32) sweep_grid<32, 64, 4, 1>(h, a);
    else if (h->cols == 16) sweep_grid<16, 64, 6, 1>(h, a);
    else if (h->tile_rows == 32) sweep_grid<64, 32, 4, 1>(h, a);
    else if (h->warps == 2) sweep_grid<64, 64, 3, 2>(h, a);
    else sweep_grid<64, 64, 3, 1>(h, a);
}

}  // namespace

// cols: 64 / 32 / 16 (0: alloc_ctx's own choice); tile_rows 64 or 32 (cols = 64 only); warps 1 or 2 (cols = 64, rows = 64 only);
// ll != 0: grad_t travels as flagged lines through the handle's own exchange block (what an attached rank does), else plain
static EmuEll* new_slab(long long n, long long rb, long long re, double kappa, const double* mq, const double* diag, const double* xc,
                        int no_defer, int use_parallel, int reduce_mode, int cols, int tile_rows, int warps, int ll) {
    EmuEll* h = new EmuEll();
    h->n = n; h->w = re - rb; h->c0 = rb;
    h->reduce_mode = reduce_mode; h->nseg = reduce_mode > 0 ? reduce_mode : 1;
    h->no_defer = no_defer; h->use_parallel = use_parallel; h->nranks = ll ? 1 : 0; h->order = 0;
    h->tile_rows = kRowsDefault; h->stages = kStagesDefault; h->warps = warps == 2 ? 2 : 1;
    // ---- alloc_ctx
    h->ld = (h->w + 1) & ~1LL;
    h->npad = (n + kRowsMax - 1) / kRowsMax * kRowsMax + kRowsMax;
    h->cols = kColsMax;
    while (h->cols > 16 && ((h->w + h->cols - 1) / h->cols) * h->nseg < 192) h->cols /= 2;
    if (cols == 16 || cols == 32 || cols == 64) h->cols = cols;
    if (h->cols != kColsMax) { h->tile_rows = 64; h->stages = h->cols == 32 ? 4 : 6; }
    else if (tile_rows == 32) { h->tile_rows = 32; h->stages = 4; }
    h->ld_part = (h->w + kColsMax - 1) / kColsMax * kColsMax;
    h->ncb = (h->w + h->cols - 1) / h->cols;
    h->S.assign((size_t)n * h->ld, 0.0);
    h->xc.assign(xc, xc + n);
    h->gpad.assign((size_t)h->npad, 0.0);
    h->gzero.assign((size_t)h->npad, 0.0);
    h->xblk.assign(xblk_doubles(n), 0.0);
    h->pend.assign((size_t)(h->npad + kColsMax + 2), make_double2(0.0, 0.0));
    h->partial.assign((size_t)h->nseg * h->ld_part, 0.0);
    h->counters.assign((size_t)(h->ncb + 2), 0u);
    emu_make_map(&h->map, h->S.data(), n, h->w, h->ld, h->tile_rows, h->cols);
    // ---- ellcuda_create: the matrix (symmetric starts only here), the scalars
    for (long long i = 0; i < n; ++i)      // the slab keeps columns [c0, c0 + w) of the (symmetric) matrix: S[i][c] = M[i][c0 + c]
        for (long long c = 0; c < h->w; ++c) {
            const long long j = h->c0 + c;
            h->S[(size_t)i * h->ld + c] = mq ? mq[i * n + j] : (i == j ? (diag ? diag[i] : 1.0) : 0.0);
        }
    std::memset(&h->sc, 0, sizeof h->sc);
    h->sc.kappa = kappa; h->sc.scale = 1.0; h->sc.epoch = 1;
    return h;
}

extern "C" void* emu_ell_new(long long n, double kappa, const double* mq, const double* diag, const double* xc, int no_defer,
                             int use_parallel, int reduce_mode, int cols, int tile_rows, int warps, int ll) {
    return new_slab(n, 0, n, kappa, mq, diag, xc, no_defer, use_parallel, reduce_mode, cols, tile_rows, warps, ll);
}

// ---- a row-sharded group in one process: rank r owns rows [r*n/P, (r+1)*n/P) (ellcuda_create with row_begin / row_end +
// ellcuda_exchange_attach_ptrs).  Every rank's sweep stores its slab of grad_t as flagged lines into EVERY rank's exchange
// block; every rank's finish polls its own block and derives the same scalars.  The ranks' kernels run one after the other.
struct EmuGroup {
    std::vector<EmuEll*> ranks;
};
extern "C" void* emu_group_new(int nranks, long long n, double kappa, const double* mq, const double* diag, const double* xc, int no_defer,
                               int use_parallel, int reduce_mode, int cols) {
    EmuGroup* g = new EmuGroup();
    for (int r = 0; r < nranks; ++r)
        g->ranks.push_back(new_slab(n, n * r / nranks, n * (r + 1) / nranks, kappa, mq, diag, xc, no_defer, use_parallel, reduce_mode, cols, 64, 1, 1));
    for (int r = 0; r < nranks; ++r) {
        g->ranks[r]->nranks = nranks;
        g->ranks[r]->rank = r;
        g->ranks[r]->peers = g->ranks;
    }
    return g;
}
static int finish_one(EmuEll* h, int which, const Cut& cut);
extern "C" int emu_group_update(void* gp, int which, const double* grad, double beta0, int has_beta1, double beta1) {
    EmuGroup* g = static_cast<EmuGroup*>(gp);
    const Cut cut = has_beta1 > 0 ? parallel_cut(beta0, beta1) : single_cut(beta0);
    for (EmuEll* h : g->ranks) {
        std::copy(grad, grad + h->n, h->gpad.begin());
        run_sweep(h, h->gpad.data(), nullptr);
    }
    int status = -1;
    for (EmuEll* h : g->ranks) {
        const int st = finish_one(h, which, cut);
        if (status >= 0 && st != status) return -100;    // the ranks must agree
        status = st;
    }
    return status;
}
// returns 0 when xc, kappa and tsq are identical on every rank (they must be), and rank 0's copy; mq = the slabs side by side
extern "C" int emu_group_state(void* gp, double* xc, double* mq, double* kappa, double* tsq) {
    EmuGroup* g = static_cast<EmuGroup*>(gp);
    int differ = 0;
    for (EmuEll* h : g->ranks) {
        run_sweep(h, nullptr, nullptr);                  // flush_pending
        h->sc.pending = 0; h->sc.scale = 1.0;
        for (long long i = 0; i < h->n; ++i)
            for (long long c = 0; c < h->w; ++c) mq[i * h->n + h->c0 + c] = h->S[(size_t)i * h->ld + c];
        const EmuEll* h0 = g->ranks[0];
        if (std::memcmp(h->xc.data(), h0->xc.data(), (size_t)h->n * 8) != 0 || std::memcmp(&h->sc.kappa, &h0->sc.kappa, 8) != 0 ||
            std::memcmp(&h->sc.tsq, &h0->sc.tsq, 8) != 0)
            differ = 1;
    }
    std::copy(g->ranks[0]->xc.begin(), g->ranks[0]->xc.end(), xc);
    *kappa = g->ranks[0]->sc.kappa;
    *tsq = g->ranks[0]->sc.tsq;
    return differ;
}
extern "C" void emu_group_free(void* gp) {
    EmuGroup* g = static_cast<EmuGroup*>(gp);
    for (EmuEll* h : g->ranks) delete h;
    delete g;
}

// ellcuda_update_begin + ellcuda_update_finish; returns the CutStatus
extern "C" int emu_ell_update(void* hp, int which, const double* g, double beta0, int has_beta1, double beta1) {
    EmuEll* h = static_cast<EmuEll*>(hp);
    std::copy(g, g + h->n, h->gpad.begin());                      // the padded copy of the gradient (copy_pad_kernel)
    run_sweep(h, h->gpad.data(), nullptr);
    const Cut cut = has_beta1 > 0 ? parallel_cut(beta0, beta1) : single_cut(beta0);
    return finish_one(h, which, cut);
}
static int finish_one(EmuEll* h, int which, const Cut& cut) {      // ellcuda_update_finish
    const CalcConsts cc = CalcConsts::make((int)h->n, h->use_parallel);
    long long fin_ctas = (h->n + kFinThreads - 1) / kFinThreads;
    if (fin_ctas > 16) fin_ctas = 16;
    for (unsigned b = 0; b < (unsigned)fin_ctas; ++b)
        emu::run_cta(kFinThreads, [&](int) {
            finish_kernel(h->gpad.data(), h->xblk.data(), h->xc.data(), h->pend.data(), &h->sc, h->counters.data() + h->ncb + 1, h->n, cc,
                          which, cut, h->no_defer, h->reduce_mode, h->nranks);
        }, EmuDim3{b, 0, 0}, EmuDim3{(unsigned)fin_ctas, 1, 1});
    return h->sc.status;
}

// Arr::dot_mv on the handle's matrix (ellcuda_dot_mv): flush first, then a sweep in plain-output mode
extern "C" void emu_ell_dot_mv(void* hp, const double* x, double* y) {
    EmuEll* h = static_cast<EmuEll*>(hp);
    run_sweep(h, nullptr, nullptr);                               // flush_pending
    h->sc.pending = 0; h->sc.scale = 1.0;                         // clear_pending_kernel
    std::copy(x, x + h->n, h->gpad.begin());
    run_sweep(h, h->gpad.data(), y);
}

// xc, the matrix as the reference holds it (pending rank-1 applied: flush_pending), kappa, tsq
extern "C" void emu_ell_state(void* hp, double* xc, double* mq, double* kappa, double* tsq) {
    EmuEll* h = static_cast<EmuEll*>(hp);
    run_sweep(h, nullptr, nullptr);
    h->sc.pending = 0; h->sc.scale = 1.0;
    std::copy(h->xc.begin(), h->xc.end(), xc);
    for (long long i = 0; i < h->n; ++i)
        for (long long j = 0; j < h->n; ++j) mq[i * h->n + j] = h->S[(size_t)i * h->ld + j];
    *kappa = h->sc.kappa;
    *tsq = h->sc.tsq;
}

extern "C" void emu_ell_free(void* hp) { delete static_cast<EmuEll*>(hp); }

#ifdef EMU_MAIN
// ThreadSanitizer driver: a few updates on small handles that walk the variants (block widths, two warps, segments with the
// cross-CTA combine, no_defer_trick, flagged lines, the ordered and the tree sum in finish), a failing cut among them.
int main() {
    struct V { int n, reduce_mode, cols, rows, warps, ll, no_defer; };
    const V variants[] = {{130, 2, 64, 64, 2, 1, 0}, {90, 3, 16, 64, 1, 0, 0}, {150, 0, 64, 64, 1, 0, 0}, {100, 0, 32, 64, 1, 1, 1}, {140, 0, 64, 32, 1, 0, 1}};
    for (const V& v : variants) {
        std::vector<double> xc(v.n), g(v.n);
        for (int i = 0; i < v.n; ++i) xc[i] = std::cos(0.3 * i);
        void* h = emu_ell_new(v.n, 2.5, nullptr, nullptr, xc.data(), v.no_defer, 1, v.reduce_mode, v.cols, v.rows, v.warps, v.ll);
        const struct { int which; double b0; int has; double b1; } cuts[] = {{0, 0.01, 1, 0.10}, {1, 0.0, -1, 0.0}, {0, 1e6, -1, 0.0}, {0, 0.02, -1, 0.0}};
        int step = 0;
        for (const auto& c : cuts) {
            for (int i = 0; i < v.n; ++i) g[i] = std::sin(0.7 * i + step) / std::sqrt((double)v.n);
            const int st = emu_ell_update(h, c.which, g.data(), c.b0, c.has, c.b1);
            std::printf("n=%d mode=%d cols=%d rows=%d warps=%d ll=%d no_defer=%d step %d: status %d\n", v.n, v.reduce_mode, v.cols, v.rows, v.warps, v.ll,
                        v.no_defer, step++, st);
        }
        std::vector<double> mq((size_t)v.n * v.n);
        double kappa = 0, tsq = 0;
        emu_ell_state(h, xc.data(), mq.data(), &kappa, &tsq);
        std::printf("  kappa %.17g tsq %.17g m00 %.17g\n", kappa, tsq, mq[0]);
        emu_ell_free(h);
    }
    return 0;
}
#endif
