// ellcta_device.cuh — device code of the CTA-per-problem batched `Ell` kernel (batch_ellcta.cu holds the description, the
// __global__ entry and the launcher).  Kept in a header so that tests/cpp/emu_ellcta.cpp can compile the very same code for
// the host — one OS thread per CUDA thread, pthread barriers for __syncthreads / ballots / shuffles — and run it under
// ThreadSanitizer: a shared-memory access that is not ordered by a barrier shows up there as a data race
// (compute-sanitizer's racecheck is not available on the GPU pool).
#pragma once
#include "common.cuh"

namespace ellcuda {
namespace ellcta {


constexpr int kCtaNMax = 224;
constexpr int kCtaWarps = kCtaNMax / 32;

struct CtaShared {
    unsigned long long pid;
    // written by warp 0 in Ell::update_core, read by everyone after the barrier
    int status;
    double tsq, rho_over_omega, ratio, delta;
    double f;                      // bowl objective
    // lowpass
    unsigned mask[2][kCtaWarps];   // per-warp violation ballots, double-buffered by round parity
    int k;
    double val;
    double bv[kCtaWarps];          // stop-band maximum per warp: value, traversal position, row
    int bc[kCtaWarps];
    int bk[kCtaWarps];
};

struct Ctx {
    const BatchArgs* a;
    CtaShared* sh;
    int tid, lane, warp, nwarps, n;
    bool on;        // this thread owns a row
    double* P;      // P[i*(i+1)/2 + j] = M[i][j], j <= i
    double* sx;     // xc (oracle input)
    double* sg;     // g
    double* srg;    // ratio * grad_t
    double* sp;     // g .* grad_t / oracle scratch
};

// Ell::new_with_matrix / new / new_with_scalar (src/ell.rs:31-78): the lower triangle
__device__ void cta_init(const Ctx& c, long long pid) {
    const BatchArgs& a = *c.a;
    const int n = c.n, t = c.tid;
    if (a.mq0) {
        const double* src = a.mq0 + pid * a.mq0_stride;
        int ti = 0;
        for (int i = 0; i < n; ++i) {
            if (t <= i) c.P[ti + t] = src[(long long)i * n + t];
            ti += i + 1;
        }
    } else {
        const int tri = n * (n + 1) / 2, nt = c.nwarps * 32;
        for (int e = t; e < tri; e += nt) c.P[e] = 0.0;
        __syncthreads();
        if (c.on) c.P[t * (t + 1) / 2 + t] = a.diag0 ? a.diag0[pid * a.diag0_stride + t] : 1.0;
    }
}

// Ell::update_core (src/ell.rs:97-137).  g = this thread's entry of the gradient (0 beyond n).  `rows_src`: the problem's
// new_with_matrix input while no cut has succeeded yet (iteration 0), else nullptr — M.g then uses the rows as given
// (:102), and the first successful cut leaves the matrix symmetric (:117-128).  Returns a status uniform over the CTA.
__device__ int cta_update(const Ctx& c, const CalcConsts& cc, int which, const Cut& cut, double g, double& xc, double& kappa,
                          double& tsq, bool no_defer, const double* rows_src) {
    const int n = c.n, t = c.tid;
    c.sg[t] = g;
    __syncthreads();
    double gt = 0.0;               // grad_t[t] = sum_j M[t][j]*g[j], j ascending from 0.0 (src/arr.rs:426-442)
    if (c.on) {
        if (rows_src) {
            const double* row = rows_src + (long long)t * n;
#pragma unroll 8
            for (int j = 0; j < n; ++j) gt += row[j] * c.sg[j];
        } else {
            const int tt = t * (t + 1) / 2;
            int tj = t;            // T(j) + t
#pragma unroll 8
            for (int j = 0; j < n; ++j) {
                gt += c.P[j <= t ? tt + j : tj] * c.sg[j];
                tj += j + 1;
            }
        }
    }
    c.sp[t] = g * gt;
    __syncthreads();
    if (c.warp == 0) {
        double omega = 0.0;        // grad.dot(grad_t), k ascending (src/arr.rs:443-451)
#pragma unroll 8
        for (int k = 0; k < n; ++k) omega += c.sp[k];
        const double tsq0 = kappa * omega;
        Rsd r;
        const int st = ell_calc(cc, which, cut, tsq0, r);
        if (c.lane == 0) {
            c.sh->status = st;
            c.sh->tsq = tsq0;
            if (st == kSuccess) {
                c.sh->rho_over_omega = r.rho / omega;
                c.sh->ratio = r.sigma / omega;
                c.sh->delta = r.delta;
            }
        }
    }
    __syncthreads();
    const int status = c.sh->status;
    tsq = c.sh->tsq;
    if (status != kSuccess) return status;
    xc -= c.sh->rho_over_omega * gt;
    c.srg[t] = c.sh->ratio * gt;
    c.sp[t] = gt;                  // warp 0 has finished reading g .* grad_t: sp now carries grad_t for the rank-1 below
    kappa *= c.sh->delta;
    __syncthreads();
    const double scale = kappa;
    if (no_defer) kappa = 1.0;     // mq *= kappa; kappa = 1 (:132-135)
    // element (i, j), j <= i: M[i][j] -= (ratio*grad_t[i]) * grad_t[j].  Column chunk ch (columns 32ch .. 32ch+31) has rows
    // 32ch .. n-1, so a warp that kept to its own columns would walk n rows in chunk 0 and n - 32ch in the others; the flat
    // list of (chunk, row) warp-steps is dealt out in equal contiguous shares instead (grad_t of a foreign chunk: one load
    // per chunk from sp, which holds grad_t since the barrier above).
    int total = 0;
    for (int ch = 0; ch < c.nwarps; ++ch) total += n - 32 * ch;
    const int share = (total + c.nwarps - 1) / c.nwarps;
    int f = c.warp * share;
    const int fend = min(total, f + share);
    int ch = 0, off = 0;
    while (ch < c.nwarps && f >= off + (n - 32 * ch)) { off += n - 32 * ch; ++ch; }
    while (f < fend) {
        const int col = 32 * ch + c.lane;
        const double gtc = c.sp[col];
        int i = 32 * ch + (f - off);
        const int iend = min(n, i + (fend - f));
        int ti = i * (i + 1) / 2;
        f += iend - i;
        if (no_defer) {
            for (; i < iend; ++i) {
                if (col <= i) c.P[ti + col] = (c.P[ti + col] - c.srg[i] * gtc) * scale;
                ti += i + 1;
            }
        } else {
#pragma unroll 4
            for (; i < iend; ++i) {
                if (col <= i) c.P[ti + col] -= c.srg[i] * gtc;
                ti += i + 1;
            }
        }
        off += n - 32 * ch;
        ++ch;
    }
    return status;
}

// ---------------------------------------------------------------------------------------------- oracles
enum : int { kCut = 0, kShrunk = 1 };

struct OrBowlCta {  // tests/integration_test.rs:90-105
    double tgt;
    __device__ void init(const Ctx& c, long long pid) {
        tgt = c.on ? c.a->params[pid * c.a->params_stride + c.tid] : 0.0;
    }
    __device__ int assess(const Ctx& c, double xc, double& gamma, double& g, Cut& cut) {
        const double d = xc - tgt;
        g = 2.0 * d;
        c.sp[c.tid] = d * d;
        __syncthreads();
        if (c.warp == 0) {
            double f0 = 0.0;
#pragma unroll 8
            for (int k = 0; k < c.n; ++k) f0 += c.sp[k];
            if (c.lane == 0) c.sh->f = f0;
        }
        __syncthreads();
        const double f = c.sh->f;
        cut = single_cut(f);
        if (f < gamma) { gamma = f; return kShrunk; }
        return kCut;
    }
};

struct OrLowpassCta {  // src/oracles/lowpass_oracle.rs:7-150; spectrum read transposed (tt[j*mdim + row]) through L1/L2
    int idx1, idx2, idx3, nwpass, nwstop, mdim, kmax, par;
    double lp_sq, up_sq, sp_sq, fmax;
    const double* tt;

    __device__ void init(const Ctx& c, long long pid) {  // :23-52
        const double* p = c.a->params + pid * c.a->params_stride;
        mdim = c.a->aux[0];
        tt = c.a->table_t;
        nwpass = (int)floor(p[0] * (double)(mdim - 1)) + 1;
        nwstop = (int)floor(p[1] * (double)(mdim - 1)) + 1;
        lp_sq = p[2]; up_sq = p[3]; sp_sq = p[4];
        idx1 = -1; idx2 = nwpass - 1; idx3 = nwstop - 1;
        fmax = -INFINITY; kmax = -1;
        par = 0;
    }
    __device__ __forceinline__ double row_elem(const Ctx& c, int k, double sign) const {
        return c.on ? sign * __ldg(c.a->table + (long long)k * c.n + c.tid) : 0.0;   // spectrum[k][t]
    }
    // one round-robin band sweep in rounds of blockDim rows (thread t takes traversal position base + t), stopping at the
    // first row in traversal order that violates.  BAND: 0 pass-band (:62-78), 1 stop-band (:83-104), 2 transition (:106-119).
    template <int BAND>
    __device__ __forceinline__ bool sweep(const Ctx& c, int lo, int hi, int& cursor, double& g, Cut& cut) {
        const int len = hi - lo, nt = c.nwarps * 32;
        double best_v = -INFINITY;
        int best_c = 0x7fffffff, best_k = -1;
        for (int base = 0; base < len; base += nt) {
            const int cidx = base + c.tid;
            int kk = cursor + 1 + cidx;
            if (kk >= hi) kk -= len;
            const bool active = cidx < len;
            const int k = active ? kk : lo;           // inactive threads read a valid row and are masked below
            double val = 0.0;                         // spectrum[k].dot(x), j ascending (src/arr.rs:443-451)
            const double* col = tt + k;
#pragma unroll 16   // 32: 180 registers, one resident CTA fewer, 10-30 % slower
            for (int j = 0; j < c.n; ++j) val += __ldg(col + (long long)j * mdim) * c.sx[j];
            bool viol;
            if (BAND == 0) viol = active && (val > up_sq || val < lp_sq);
            else if (BAND == 1) viol = active && (val > sp_sq || val < 0.0);
            else viol = active && (val < 0.0);
            const unsigned mask = __ballot_sync(0xffffffffu, viol);
            unsigned* mk = c.sh->mask[par];
            par ^= 1;
            if (c.lane == 0) mk[c.warp] = mask;
            __syncthreads();
            int fw = -1;
            unsigned fm = 0;
            for (int w = c.nwarps - 1; w >= 0; --w) {
                const unsigned m = mk[w];
                if (m) { fw = w; fm = m; }
            }
            if (fw >= 0) {                            // uniform over the CTA
                const int ft = fw * 32 + __ffs(fm) - 1;
                if (c.tid == ft) { c.sh->k = k; c.sh->val = val; }
                __syncthreads();
                const int kf = c.sh->k;
                const double vf = c.sh->val;
                cursor = kf;
                if (BAND == 0) {
                    if (vf > up_sq) { g = row_elem(c, kf, 1.0); cut = parallel_cut(vf - up_sq, vf - lp_sq); }
                    else { g = row_elem(c, kf, -1.0); cut = parallel_cut(-vf + lp_sq, -vf + up_sq); }
                } else if (BAND == 1) {
                    if (vf > sp_sq) { g = row_elem(c, kf, 1.0); cut = parallel_cut(vf - sp_sq, vf); }
                    else { g = row_elem(c, kf, -1.0); cut = parallel_cut(-vf, -vf + sp_sq); }
                } else {
                    g = row_elem(c, kf, -1.0);
                    cut = single_cut(-vf);            // ParallelCut(-val, None)
                }
                return true;
            }
            if (BAND == 1 && active && val > best_v) { best_v = val; best_c = cidx; best_k = k; }   // first occurrence of the maximum
        }
        int last = cursor + len;                       // no violation: the cursor sits on the last row visited
        if (last >= hi) last -= len;
        cursor = last;
        if (BAND == 1) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, best_v, off);
                const int oc = __shfl_xor_sync(0xffffffffu, best_c, off);
                const int ok = __shfl_xor_sync(0xffffffffu, best_k, off);
                if (ov > best_v || (ov == best_v && oc < best_c)) { best_v = ov; best_c = oc; best_k = ok; }
            }
            if (c.lane == 0) { c.sh->bv[c.warp] = best_v; c.sh->bc[c.warp] = best_c; c.sh->bk[c.warp] = best_k; }
            __syncthreads();
            best_v = c.sh->bv[0]; best_c = c.sh->bc[0]; best_k = c.sh->bk[0];
            for (int w = 1; w < c.nwarps; ++w) {
                const double ov = c.sh->bv[w];
                const int oc = c.sh->bc[w], ok = c.sh->bk[w];
                if (ov > best_v || (ov == best_v && oc < best_c)) { best_v = ov; best_c = oc; best_k = ok; }
            }
            fmax = best_v;
            kmax = best_k;
        }
        return false;
    }
    __device__ int assess(const Ctx& c, double xc, double& gamma, double& g, Cut& cut) {  // :137-150
        (void)xc;
        sp_sq = gamma;
        if (sweep<0>(c, 0, nwpass, idx1, g, cut)) return kCut;
        fmax = -INFINITY; kmax = -1;
        if (sweep<1>(c, nwstop, mdim, idx3, g, cut)) return kCut;
        if (sweep<2>(c, nwpass, nwstop, idx2, g, cut)) return kCut;
        const double x0 = c.sx[0];
        if (x0 < 0.0) {  // :123-127
            g = c.tid == 0 ? -1.0 : 0.0;
            cut = single_cut(-x0);
            return kCut;
        }
        g = row_elem(c, kmax, 1.0);
        cut = parallel_cut(0.0, fmax);
        gamma = fmax;
        return kShrunk;
    }
};

// ---------------------------------------------------------------------------------------------- kernel
// the whole kernel as seen by thread `tid` of a CTA of `nt` threads; `smem` = the CTA's dynamic shared memory, `sh` = its static part
template <class Oracle>
__device__ void ellcta_body(const BatchArgs& a, double* smem, CtaShared& sh, int tid, int nt) {
    const int n = a.n;
    Ctx c;
    c.a = &a; c.sh = &sh;
    c.tid = tid; c.lane = tid & 31; c.warp = tid >> 5; c.nwarps = nt >> 5;
    c.n = n; c.on = c.tid < n;
    c.P = smem;
    double* vec = smem + ((n * (n + 1) / 2 + 1) & ~1);
    c.sx = vec; c.sg = vec + nt; c.srg = vec + 2 * nt; c.sp = vec + 3 * nt;
    const CalcConsts cc = CalcConsts::make(n, a.use_parallel_cut);
    const int t = c.tid;
    Oracle om;

    for (;;) {
        __syncthreads();                              // everyone has read the previous problem's pid and shared state
        if (t == 0) sh.pid = atomicAdd(a.queue, 1ULL);
        __syncthreads();
        const long long pid = (long long)sh.pid;
        if (pid >= a.count) break;

        cta_init(c, pid);
        double xc = c.on ? a.xc0[pid * a.xc0_stride + t] : 0.0;
        double kappa = a.kappa0[pid * a.kappa0_stride];
        double gamma = a.gamma0 ? a.gamma0[pid * a.gamma0_stride] : 0.0;
        double tsq = 0.0;
        om.init(c, pid);
        bool has_best = false;
        double best = 0.0;
        long long niter = 0;
        int exit_status = kExitMaxIt;

        for (; niter < a.max_iters; ++niter) {   // cutting_plane_optim :299-311
            c.sx[t] = xc;
            __syncthreads();
            double g = 0.0;
            Cut cut = single_cut(0.0);
            const int what = om.assess(c, xc, gamma, g, cut);
            if (what == kShrunk) { has_best = true; best = xc; }
            if (!c.on) g = 0.0;
            const double* rows = (a.mq0 != nullptr && niter == 0) ? a.mq0 + pid * a.mq0_stride : nullptr;
            const int status = cta_update(c, cc, what == kShrunk ? kCentral : kBias, cut, g, xc, kappa, tsq,
                                          a.no_defer_trick != 0, rows);
            if (status != kSuccess) { exit_status = status; break; }
            if (tsq < a.tolerance) { exit_status = kExitTol; break; }
        }
        if (t == 0) {
            a.num_iters[pid] = niter;
            a.exit_status[pid] = exit_status;
            if (a.has_xbest) a.has_xbest[pid] = has_best ? 1 : 0;
            if (a.gamma) a.gamma[pid] = gamma;
            if (a.tsq) a.tsq[pid] = tsq;
            if (a.kappa) a.kappa[pid] = kappa;
        }
        if (c.on) {
            if (a.xbest && has_best) a.xbest[pid * n + t] = best;
            if (a.xc) a.xc[pid * n + t] = xc;
        }
    }
}


}  // namespace ellcta
}  // namespace ellcuda
