// emu_sym.cpp — the large-n path with SYMMETRIC STORAGE (ellalgo_rs_b200/csrc/stream_sym.cu: the lower block-triangle of M as
// 64x64 tiles; sym_sweep_kernel with its two passes over a tile, the ordered sym_combine_kernel; sharded over ranks:
// sym_combine_push_kernel + sym_reduce_kernel, a reduce-scatter and an all-gather of flagged lines) compiled for the host with
// cuda_emu.h + stream_ptx_emu.h, followed by stream_update.cu's finish_kernel.  Both .cu files are included as they are (host
// sides behind #ifdef __CUDACC__); the handle's bookkeeping (sym_alloc / sym_upload / sym_sweep / sym_reduce / sym_unpack) is
// restated for host memory, the grids' CTAs — and, for a sharded group, the ranks — run one after the other.
// tests/test_emulated_kernels.py: within 1e-9 of the CPU checker (this storage regroups the sums; the device path is held to
// the same bar), identical on every rank of a group, race-free under ThreadSanitizer.  Test infrastructure only.
#include "cuda_emu.h"
#include "stream_ptx_emu.h"

#include "../../ellalgo_rs_b200/csrc/stream_update.cu"
#include "../../ellalgo_rs_b200/csrc/stream_sym.cu"

namespace ellcuda {
void set_error(const std::string&) {}
void keep_mempool(int) {}
namespace {
alignas(128) unsigned char smem_raw[160 * 1024];   // the sweep CTA's dynamic shared memory
}
}  // namespace ellcuda

using namespace ellcuda;

namespace {

struct EmuSym {
    long long n, T, npad, slice, gpad_n;
    int nseg, rank, count, reduce_mode, no_defer, use_parallel, use_map;
    std::vector<double> Mp, colpart, rowpart, xc, gpad, gzero, xblk;
    std::vector<double2> pend;
    unsigned arrivals = 0;
    Scal sc;
    CUtensorMap map;
    std::vector<EmuSym*> peers;
};

EmuSym* new_sym(long long n, int rank, int count, double kappa, const double* mq, const double* diag, const double* xc, int no_defer,
                int use_parallel, int reduce_mode, int use_map) {
    EmuSym* h = new EmuSym();
    h->n = n; h->rank = rank; h->count = count; h->reduce_mode = reduce_mode; h->no_defer = no_defer; h->use_parallel = use_parallel;
    h->use_map = use_map;
    // ---- sym_alloc
    h->T = (n + kTile - 1) / kTile;
    h->npad = h->T * kTile;
    const long long all_pairs = (h->T + 1) / 2;
    long long npairs = (all_pairs - rank + count - 1) / count;
    if (npairs < 1) npairs = 1;
    h->slice = (n + count - 1) / count;
    int nseg = reduce_mode > 0 ? reduce_mode : (int)((256 + npairs - 1) / npairs);
    if (nseg < 1) nseg = 1;
    if (nseg > 16) nseg = 16;
    h->nseg = nseg;
    const size_t tiles = (size_t)tile_index(h->T - 1, h->T - 1) + 1;
    h->Mp.assign(tiles * kTileDoubles, 0.0);
    h->colpart.assign((size_t)nseg * h->npad, 0.0);
    h->rowpart.assign((size_t)h->T * h->npad, 0.0);
    emu_make_map(&h->map, h->Mp.data(), (long long)tiles * kTile, kTile, kTile, kTile, kTile);
    // ---- alloc_ctx (the buffers shared with the full-storage path)
    h->gpad_n = (n + kRowsMax - 1) / kRowsMax * kRowsMax + kRowsMax;
    h->xc.assign(xc, xc + n);
    h->gpad.assign((size_t)h->gpad_n, 0.0);
    h->gzero.assign((size_t)h->gpad_n, 0.0);
    h->pend.assign((size_t)(h->gpad_n + kColsMax + 2), make_double2(0.0, 0.0));
    h->xblk.assign(xblk_doubles(n) + (count > 1 ? (size_t)4 * count * h->slice : 0), 0.0);
    // ---- sym_upload
    if (!mq) {
        for (long long i = 0; i < n; ++i) h->Mp[(size_t)tile_index(i >> 6, i >> 6) * kTileDoubles + (i & 63) * kTile + (i & 63)] = diag ? diag[i] : 1.0;
    } else {
        for (long long I = 0; I < h->T; ++I)
            for (long long J = 0; J <= I; ++J)
                for (long long r = 0; r < kTile; ++r) {
                    const long long i = I * kTile + r;
                    if (i >= n) break;
                    const long long cmax = std::min<long long>(kTile, n - J * kTile);
                    std::memcpy(&h->Mp[(size_t)tile_index(I, J) * kTileDoubles + r * kTile], mq + i * n + J * kTile, (size_t)cmax * 8);
                }
    }
    std::memset(&h->sc, 0, sizeof h->sc);
    h->sc.kappa = kappa; h->sc.scale = 1.0; h->sc.epoch = 1;
    return h;
}

PeerTable peer_table(const EmuSym* h) {
    PeerTable pt;
    pt.nranks = h->count > 1 ? h->count : 0;
    pt.rank = h->rank;
    for (int r = 0; r < kMaxRanks; ++r) pt.blk[r] = r < pt.nranks ? h->peers[r]->xblk.data() : nullptr;
    return pt;
}

void grid_1d(unsigned blocks, int threads, const std::function<void()>& body) {
    for (unsigned b = 0; b < blocks; ++b) emu::run_cta(threads, [&](int) { body(); }, EmuDim3{b, 0, 0}, EmuDim3{blocks, 1, 1});
}

// sym_sweep: the tile sweep, then the combine (single handle) or the push to the reducing ranks (sharded)
void sweep(EmuSym* h, const double* g, int exchange) {
    SymArgs a;
    a.Mp = h->Mp.data(); a.g = g; a.pend = h->pend.data(); a.sc = &h->sc; a.colpart = h->colpart.data(); a.rowpart = h->rowpart.data();
    a.npad = h->npad; a.T = (int)h->T; a.nseg = h->nseg; a.exchange = exchange; a.use_map = h->use_map;
    a.rank = h->rank; a.count = h->count;
    const long long all_pairs = (h->T + 1) / 2;
    const long long own_pairs = all_pairs > h->rank ? (all_pairs - h->rank + h->count - 1) / h->count : 0;
    for (unsigned seg = 0; seg < (unsigned)h->nseg && own_pairs > 0; ++seg)
        for (unsigned p = 0; p < (unsigned)own_pairs; ++p) {
            std::memset(smem_raw, 0xff, sizeof smem_raw);
            emu::run_cta(64, [&](int) { sym_sweep_kernel(h->map, a); }, EmuDim3{p, seg, 0}, EmuDim3{(unsigned)own_pairs, (unsigned)h->nseg, 1});
        }
    if (!exchange) return;
    const unsigned blocks = (unsigned)((h->n + 31) / 32);
    if (h->count > 1) {
        const PeerTable pt = peer_table(h);
        grid_1d(blocks, kCombThreads, [&] { sym_combine_push_kernel<false>(h->colpart.data(), h->rowpart.data(), h->n, a.npad, a.nseg, pt, &h->sc, h->slice); });
    } else {
        grid_1d(blocks, kCombThreads, [&] { sym_combine_kernel(h->colpart.data(), h->rowpart.data(), h->xblk.data(), h->n, a.npad, a.nseg); });
    }
}

void reduce(EmuSym* h) {   // sym_reduce (sharded): this rank's slice, the P partials in rank order, broadcast as flagged lines
    const PeerTable pt = peer_table(h);
    grid_1d((unsigned)((h->slice + 255) / 256), 256, [&] { sym_reduce_kernel(h->n, pt, &h->sc, h->slice); });
}

int finish(EmuSym* h, int which, const Cut& cut) {
    const CalcConsts cc = CalcConsts::make((int)h->n, h->use_parallel);
    long long fin_ctas = (h->n + kFinThreads - 1) / kFinThreads;
    if (fin_ctas > 16) fin_ctas = 16;
    grid_1d((unsigned)fin_ctas, kFinThreads, [&] {
        finish_kernel(h->gpad.data(), h->xblk.data(), h->xc.data(), h->pend.data(), &h->sc, &h->arrivals, h->n, cc, which, cut, h->no_defer,
                      h->reduce_mode == 0 ? 1 : h->reduce_mode, h->count > 1 ? h->count : 0);
    });
    return h->sc.status;
}

void unpack(EmuSym* h, double* mq) {   // flush_pending + sym_unpack
    sweep(h, h->gzero.data(), 0);
    h->sc.pending = 0; h->sc.scale = 1.0;
    const long long n = h->n;
    for (long long i = 0; i < n; ++i)
        for (long long j = 0; j < n; ++j) {
            const long long I = i >> 6, J = j >> 6, r = i & 63, c = j & 63;
            const bool mine = h->count <= 1 || (std::min(std::min(I, J), h->T - 1 - std::min(I, J)) % h->count) == h->rank;
            if (!mine) continue;       // a sharded rank holds the tiles of its column-block pairs only
            mq[i * n + j] = I >= J ? h->Mp[(size_t)tile_index(I, J) * kTileDoubles + r * kTile + c] : h->Mp[(size_t)tile_index(J, I) * kTileDoubles + c * kTile + r];
        }
}

}  // namespace

struct EmuSymGroup {
    std::vector<EmuSym*> ranks;
};

// nranks = 1: a single handle with the whole block-triangle; > 1: ELL_STORAGE_SYMMETRIC sharded by shard_rank / shard_count
extern "C" void* emu_sym_new(int nranks, long long n, double kappa, const double* mq, const double* diag, const double* xc, int no_defer,
                             int use_parallel, int reduce_mode, int use_map) {
    EmuSymGroup* g = new EmuSymGroup();
    for (int r = 0; r < nranks; ++r) g->ranks.push_back(new_sym(n, r, nranks, kappa, mq, diag, xc, no_defer, use_parallel, reduce_mode, use_map));
    for (EmuSym* h : g->ranks) h->peers = g->ranks;
    return g;
}
extern "C" int emu_sym_update(void* gp, int which, const double* grad, double beta0, int has_beta1, double beta1) {
    EmuSymGroup* g = static_cast<EmuSymGroup*>(gp);
    const Cut cut = has_beta1 > 0 ? parallel_cut(beta0, beta1) : single_cut(beta0);
    for (EmuSym* h : g->ranks) {
        std::copy(grad, grad + h->n, h->gpad.begin());
        sweep(h, h->gpad.data(), 1);
    }
    if (g->ranks.size() > 1)
        for (EmuSym* h : g->ranks) reduce(h);
    int status = -1;
    for (EmuSym* h : g->ranks) {
        const int st = finish(h, which, cut);
        if (status >= 0 && st != status) return -100;    // the ranks must agree
        status = st;
    }
    return status;
}
extern "C" int emu_sym_state(void* gp, double* xc, double* mq, double* kappa, double* tsq) {
    EmuSymGroup* g = static_cast<EmuSymGroup*>(gp);
    int differ = 0;
    const EmuSym* h0 = g->ranks[0];
    for (EmuSym* h : g->ranks) {
        unpack(h, mq);
        if (std::memcmp(h->xc.data(), h0->xc.data(), (size_t)h->n * 8) != 0 || std::memcmp(&h->sc.kappa, &h0->sc.kappa, 8) != 0 ||
            std::memcmp(&h->sc.tsq, &h0->sc.tsq, 8) != 0)
            differ = 1;
    }
    std::copy(h0->xc.begin(), h0->xc.end(), xc);
    *kappa = h0->sc.kappa;
    *tsq = h0->sc.tsq;
    return differ;
}
extern "C" void emu_sym_free(void* gp) {
    EmuSymGroup* g = static_cast<EmuSymGroup*>(gp);
    for (EmuSym* h : g->ranks) delete h;
    delete g;
}

#ifdef EMU_MAIN
int main() {   // ThreadSanitizer driver: a single handle and a sharded group of three, TMA-map and plain bulk copies
    struct V { int nranks, n, use_map, no_defer; };
    const V variants[] = {{1, 130, 1, 0}, {3, 200, 1, 1}, {1, 100, 0, 0}};
    for (const V& v : variants) {
        std::vector<double> xc(v.n), g(v.n), mq((size_t)v.n * v.n);
        for (int i = 0; i < v.n; ++i) xc[i] = std::cos(0.3 * i);
        void* h = emu_sym_new(v.nranks, v.n, 2.5, nullptr, nullptr, xc.data(), v.no_defer, 1, 0, v.use_map);
        const struct { int which; double b0; int has; double b1; } cuts[] = {{0, 0.01, 1, 0.10}, {0, 1e6, -1, 0.0}, {1, 0.0, -1, 0.0}};
        int step = 0;
        for (const auto& c : cuts) {
            for (int i = 0; i < v.n; ++i) g[i] = std::sin(0.7 * i + step) / std::sqrt((double)v.n);
            std::printf("ranks=%d n=%d map=%d no_defer=%d step %d: status %d\n", v.nranks, v.n, v.use_map, v.no_defer, step, emu_sym_update(h, c.which, g.data(), c.b0, c.has, c.b1));
            ++step;
        }
        double kappa = 0, tsq = 0;
        const int differ = emu_sym_state(h, xc.data(), mq.data(), &kappa, &tsq);
        std::printf("  kappa %.17g tsq %.17g m00 %.17g ranks differ %d\n", kappa, tsq, mq[0], differ);
        emu_sym_free(h);
    }
    return 0;
}
#endif
