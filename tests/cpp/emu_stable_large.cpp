// emu_stable_large.cpp — the single-handle EllStable update for 64 < n <= 2048 (ellalgo_rs_b200/csrc/stable_large_device.cuh:
// one CTA of 1024 threads, the packed factor in global memory, the vectors in shared memory) compiled for the host with
// cuda_emu.h: kBigT OS threads (1024 as on the device; 128 in the ThreadSanitizer build, which tracks at most 256 running
// threads exactly — the loops are strided by kBigT, so the hand-offs between threads are the same in kind), a pthread barrier for __syncthreads.  tests/test_emulated_kernels.py checks it against the CPU
// checker bit for bit and runs it under ThreadSanitizer (the forward substitution, the back substitution and the factor
// update hand data between threads at every one of their ~3n barriers).  Test infrastructure only.
#include "cuda_emu.h"

#include "../../ellalgo_rs_b200/csrc/stable_large_device.cuh"

namespace ellcuda {
void set_error(const std::string&) {}
}  // namespace ellcuda

using namespace ellcuda;
using namespace ellcuda::stable_large;

// one update_core (src/ell_stable.rs:52-125) on qu = packed upper triangle incl. the diagonal (row r at r*n - r(r-1)/2),
// scal = {kappa, tsq}; has_beta1: <= 0 single cut, > 0 ParallelCut(beta0, Some(beta1)) — as ellcuda_update passes it
extern "C" int emu_stable_large_update(double* qu, double* xc, double* scal, int* status, const double* g, int n, int use_parallel,
                                       int which, double beta0, int has_beta1, double beta1) {
    if (n < 1 || n > kBigNMax) return ELLCUDA_E_ARG;
    const CalcConsts cc = CalcConsts::make(n, use_parallel);
    const Cut cut = has_beta1 > 0 ? parallel_cut(beta0, beta1) : single_cut(beta0);
    std::vector<double> sh((size_t)7 * n, NAN);     // uninitialised shared memory must never reach a result
    StableLargeShared ss;
    std::memset(&ss, 0xff, sizeof ss);
    emu::run_cta(kBigT, [&](int t) { stable_large_body(qu, xc, scal, status, g, n, cc, which, cut, sh.data(), ss, t); });
    return ELLCUDA_OK;
}

#ifdef EMU_MAIN
// ThreadSanitizer driver: a few updates (parallel cut, central cut, a failing cut) of an n = 70 factor with a non-trivial
// strict upper triangle.
int main() {
    const int n = 70;
    std::vector<double> qu((size_t)n * (n + 1) / 2), xc(n), g(n), scal{2.5, 0.0};
    for (int r = 0; r < n; ++r)
        for (int c = r; c < n; ++c)
            qu[(size_t)upper_off_ll(r, n) + (c - r)] = c == r ? 1.0 + 0.01 * r : 1e-3 * std::sin(1.0 + 3.0 * r + 7.0 * c);
    for (int i = 0; i < n; ++i) xc[i] = std::cos(0.3 * i);
    int status = -1, bad = 0;
    const struct { int which; double b0; int has; double b1; } cuts[] = {{0, 0.01, 1, 0.10}, {1, 0.0, -1, 0.0}, {0, 1e6, -1, 0.0}, {0, 0.02, -1, 0.0}};
    int step = 0;
    for (const auto& c : cuts) {
        for (int i = 0; i < n; ++i) g[i] = std::sin(0.7 * i + step) / std::sqrt((double)n);
        bad |= emu_stable_large_update(qu.data(), xc.data(), scal.data(), &status, g.data(), n, 1, c.which, c.b0, c.has, c.b1);
        std::printf("step %d: status %d, kappa %.17g, tsq %.17g\n", step, status, scal[0], scal[1]);
        ++step;
    }
    return bad;
}
#endif
