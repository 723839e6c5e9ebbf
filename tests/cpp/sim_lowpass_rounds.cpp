// Rounds of 32 spectrum rows per LowpassOracle call under the lane<->row mappings discussed in DESIGN.md 4.2b, counted on the
// checker's own trajectories: cursor-relative windows (batch_warp.cu), windows that stop at the band's wrap point
// (batch_lowpass.cu, default), block-aligned rounds (TMEM variant 2), aligned rounds with the head of the first block
// deferred, and two rows per lane.  Test infrastructure (it includes oracle/): built and run by tests/test_design_numbers.py.
#include "../../oracle/ellref.hpp"
#include "../../oracle/ellref_oracles.hpp"
#include <cstdio>
#include <cstdlib>
using namespace ellref;
static long long g_pack = 0, g_callrows = 0, g_rows = 0;   // dense packing across bands: ceil(rows visited in the call / 32)
static long long g_calls = 0, g_cur = 0, g_blk = 0, g_sweeps = 0, g_win = 0, g_r2 = 0, g_def = 0;
static void account(int lo, int hi, int start /*first row visited*/, int visited /*rows visited incl. violator, or len*/) {
    const int len = hi - lo;
    g_sweeps++; g_callrows += visited; g_rows += visited;
    g_cur += (visited + 31) / 32;
    // block-aligned: walk rows in traversal order, count segments
    int segs = 0, k = start, left = visited;
    while (left > 0) {
        int end = std::min(hi, (k / 32 + 1) * 32);
        int take = std::min(left, end - k);
        segs++; left -= take; k += take;
        if (k >= hi) k = lo;
    }
    g_blk += segs;
    { int w = 0, kk = start, l2 = visited; while (l2 > 0) { int take = std::min(l2, std::min(32, hi - kk)); w++; l2 -= take; kk += take; if (kk >= hi) kk = lo; } g_win += w; }
    g_r2 += 2 * ((visited + 63) / 64);
    {   // aligned blocks, head of the first block deferred (evaluated in the first round, consulted last)
        int w = 0, kk = start, l2 = visited; bool first = true; int head = start - std::max(lo, (start / 32) * 32);
        while (l2 > 0) {
            int end = std::min(hi, (kk / 32 + 1) * 32);
            int take = std::min(l2, end - kk);
            // the final wrapped piece [blockstart(start), start) was already evaluated in round 1
            bool is_final_head = !first && (kk / 32 == start / 32) && kk < start && kk >= std::max(lo, (start/32)*32) && head > 0;
            if (!is_final_head) w++;
            first = false; l2 -= take; kk += take; if (kk >= hi) kk = lo;
        }
        g_def += w;
    }
    (void)len;
}
struct SimLowpass : Lowpass {
    using Lowpass::Lowpass;
    FeasAnswer assess_feas(const Vec& x) {
        g_calls++;
        more_alt = true;
        { int start = idx1 + 1 == nwpass ? 0 : idx1 + 1; int v = 0;
        for (int c = 0; c < nwpass; ++c) {
            idx1 += 1; if (idx1 == nwpass) idx1 = 0; v++;
            const double val = dot(spectrum + idx1 * ndim, x.data(), ndim);
            if (val > up_sq) { account(0, nwpass, start, v); return CutPair{row(idx1, 1.0), Cut::parallel(val - up_sq, val - lp_sq)}; }
            if (val < lp_sq) { account(0, nwpass, start, v); return CutPair{row(idx1, -1.0), Cut::parallel(-val + lp_sq, -val + up_sq)}; }
        } account(0, nwpass, start, v); }
        fmax = -std::numeric_limits<double>::infinity(); kmax = -1;
        { int start = idx3 + 1 == (int)mdim ? nwstop : idx3 + 1; int v = 0;
        for (int c = nwstop; c < static_cast<int>(mdim); ++c) {
            idx3 += 1; if (idx3 == static_cast<int>(mdim)) idx3 = nwstop; v++;
            const double val = dot(spectrum + idx3 * ndim, x.data(), ndim);
            if (val > sp_sq) { account(nwstop, mdim, start, v); return CutPair{row(idx3, 1.0), Cut::parallel(val - sp_sq, val)}; }
            if (val < 0.0) { account(nwstop, mdim, start, v); return CutPair{row(idx3, -1.0), Cut::parallel(-val, -val + sp_sq)}; }
            if (val > fmax) { fmax = val; kmax = idx3; }
        } account(nwstop, mdim, start, v); }
        { int start = idx2 + 1 == nwstop ? nwpass : idx2 + 1; int v = 0;
        for (int c = nwpass; c < nwstop; ++c) {
            idx2 += 1; if (idx2 == nwstop) idx2 = nwpass; v++;
            const double val = dot(spectrum + idx2 * ndim, x.data(), ndim);
            if (val < 0.0) { account(nwpass, nwstop, start, v); return CutPair{row(idx2, -1.0), Cut::single(-val)}; }
        } account(nwpass, nwstop, start, v); }
        more_alt = false;
        if (x[0] < 0.0) { Vec g(ndim, 0.0); g[0] = -1.0; return CutPair{g, Cut::single(-x[0])}; }
        return std::nullopt;
    }
    OptimAnswer assess_optim(const Vec& x, double& sp) {
        sp_sq = sp;
        g_callrows = 0;
        auto c = assess_feas(x);
        g_pack += (g_callrows + 31) / 32;
        if (c) return {*c, false};
        CutPair cut{row(kmax, 1.0), Cut::parallel(0.0, fmax)};
        sp = fmax;
        return {cut, true};
    }
};
int main(int argc, char** argv) {
    const int count = argc > 1 ? atoi(argv[1]) : 64;
    const std::size_t n = 32;
    Vec table = lowpass_spectrum(n);
    unsigned long long s = 12345;
    auto rnd = [&]() { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(s >> 11) / 9007199254740992.0; };
    long long iters = 0;
    for (int i = 0; i < count; ++i) {
        double wpass = 0.10 + 0.04 * rnd(), dp = 0.02 + 0.01 * rnd(), ds = 0.10 + 0.05 * rnd();
        double d1 = 20.0 * std::log10(1.0 + dp), lp = std::pow(10.0, -d1 / 20.0), up = std::pow(10.0, d1 / 20.0);
        SimLowpass o(n, table.data(), wpass, wpass + 0.08, lp * lp, up * up, ds * ds);
        Ell sp = Ell::with_scalar(40.0, Vec(n, 0.0));
        double gamma = ds * ds;
        Options opt{50000, 1e-14};
        auto rr = cutting_plane_optim(o, sp, gamma, opt);
        iters += rr.num_iters;
    }
    printf("problems %d iterations %lld oracle calls %lld sweeps %lld\n", count, iters, g_calls, g_sweeps);
    printf("rounds/call: current %.3f  block-aligned %.3f  window-break-at-wrap %.3f  R2 %.3f  aligned+deferred-head %.3f\n", (double)g_cur / g_calls, (double)g_blk / g_calls, (double)g_win / g_calls, (double)g_r2 / g_calls, (double)g_def / g_calls);
    printf("dense packing across bands (lower bound of any round-merging scheme): %.3f rounds/call, %.1f rows/call\n", (double)g_pack / g_calls, (double)g_rows / g_calls);
    return 0;
}
