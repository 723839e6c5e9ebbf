// batch_small.cu — batched cutting-plane loop for tiny problems (n = 2): ONE THREAD PER PROBLEM.
//
// At n = 2 the whole search space (4 + 2 + 2 doubles) and the oracle state fit in registers, so a problem is a
// thread; a warp pulls 32 problems at a time from a global counter (persistent grid, no host round trip per
// iteration).  The loop bodies restate
//   cutting_plane_optim / _feas / _optim_q            src/cutting_plane.rs:286-313, 205-227, 331-374
//   Ell::update_core                                  src/ell.rs:97-137
//   EllStable::update_core                            src/ell_stable.rs:52-125
// with the oracles of src/example1.rs, example1_rr.rs, example3.rs, example4.rs, quasicvx.rs,
// src/oracles/profit_oracle.rs and the fixtures of tests/cutting_plane_tests.rs, tests/example2_tests.rs.
// Compiled with -fmad=false: no contraction, sums in the reference's order.  exp/ln are glibc's, bit for bit
// (glibc_math.cuh), because that is what f64::exp / f64::ln are on the reference's platform.
#include "common.cuh"
#include "glibc_math.cuh"

namespace ellcuda {
namespace {

// ------------------------------------------------------------------------------------------------ spaces
template <int N>
struct EllReg {  // src/ell.rs:9-16
    double m[N * N];
    double xc[N];
    double kappa, tsq;
    bool no_defer;

    __device__ int update(const CalcConsts& cc, int which, const double (&g)[N], const Cut& cut) {  // :97-137
        double gt[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < N; ++j) acc += m[i * N + j] * g[j];
            gt[i] = acc;
        }
        double omega = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) omega += g[i] * gt[i];
        tsq = kappa * omega;
        Rsd r;
        const int status = ell_calc(cc, which, cut, tsq, r);
        if (status != kSuccess) return status;
        const double rho_over_omega = r.rho / omega;
#pragma unroll
        for (int i = 0; i < N; ++i) xc[i] -= rho_over_omega * gt[i];
        const double ratio = r.sigma / omega;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double r_qg = ratio * gt[i];
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const double upd = r_qg * gt[j];
                m[i * N + j] -= upd;
                if (i != j) m[j * N + i] = m[i * N + j];
            }
        }
        kappa *= r.delta;
        if (no_defer) {
#pragma unroll
            for (int k = 0; k < N * N; ++k) m[k] *= kappa;
            kappa = 1.0;
        }
        return status;
    }
};

template <int N>
struct EllStableReg {  // src/ell_stable.rs:9-15
    double m[N * N];
    double xc[N];
    double kappa, tsq;
    bool no_defer;  // unused (EllStable has no defer switch)

    __device__ int update(const CalcConsts& cc, int which, const double (&g)[N], const Cut& cut) {  // :52-125
        double w[N];
#pragma unroll
        for (int i = 0; i < N; ++i) w[i] = g[i];
#pragma unroll
        for (int i = 1; i < N; ++i) {
#pragma unroll
            for (int j = 0; j < i; ++j) {
                const double val = m[j * N + i] * w[j];
                m[i * N + j] = val;
                w[i] -= val;
            }
        }
        double z[N], gg[N];
#pragma unroll
        for (int i = 0; i < N; ++i) z[i] = w[i] * m[i * N + i];
        double omega = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            gg[i] = z[i] * w[i];
            omega += gg[i];
        }
        tsq = kappa * omega;
        Rsd r;
        const int status = ell_calc(cc, which, cut, tsq, r);
        if (status != kSuccess) return status;
        double q[N];
#pragma unroll
        for (int i = 0; i < N; ++i) q[i] = z[i];
#pragma unroll
        for (int i = N - 1; i >= 1; --i) {
#pragma unroll
            for (int j = i; j < N; ++j) q[i - 1] -= m[j * N + (i - 1)] * q[j];
        }
        const double rho_over_omega = r.rho / omega;
#pragma unroll
        for (int i = 0; i < N; ++i) xc[i] -= rho_over_omega * q[i];
        const double mu = r.sigma / (1.0 - r.sigma);
        double oldt = omega / mu;
#pragma unroll
        for (int j = 0; j < N - 1; ++j) {
            const double temp = oldt + gg[j];
            const double beta2 = z[j] / temp;
            m[j * N + j] *= oldt / temp;
#pragma unroll
            for (int l = j + 1; l < N; ++l) m[j * N + l] += beta2 * m[l * N + j];
            oldt = temp;
        }
        const double temp = oldt + gg[N - 1];
        m[(N - 1) * N + (N - 1)] *= oldt / temp;
        kappa *= r.delta;
        return status;
    }
};

// ------------------------------------------------------------------------------------------------ oracles (n = 2)
// Interface: init(args, pid);  optim: bool assess_optim(x, gamma, g, cut) -> shrunk;
//            feas: bool assess_feas(x, g, cut) -> true if a cut was returned (false = feasible);
//            optim_q: assess_optim_q(x, gamma, retry, g, cut, xq, more_alt) -> shrunk.
struct OrExample1 {  // src/example1.rs:8-30
    static constexpr int kN = 2;
    __device__ void init(const BatchArgs&, long long) {}
    __device__ bool assess_optim(const double (&x)[2], double& gamma, double (&g)[2], Cut& cut) {
        const double f0 = x[0] + x[1];
        const double f1 = f0 - 3.0;
        if (f1 > 0.0) { g[0] = 1.0; g[1] = 1.0; cut = single_cut(f1); return false; }
        const double f2 = -x[0] + x[1] + 1.0;
        if (f2 > 0.0) { g[0] = -1.0; g[1] = 1.0; cut = single_cut(f2); return false; }
        const double f3 = gamma - f0;
        if (f3 > 0.0) { g[0] = -1.0; g[1] = -1.0; cut = single_cut(f3); return false; }
        gamma = f0;
        g[0] = -1.0; g[1] = -1.0; cut = single_cut(0.0);
        return true;
    }
};

struct OrExample1RR {  // src/example1_rr.rs:13-53
    static constexpr int kN = 2;
    int idx;
    __device__ void init(const BatchArgs&, long long) { idx = -1; }
    __device__ bool assess_optim(const double (&x)[2], double& gamma, double (&g)[2], Cut& cut) {
        const double f0 = x[0] + x[1];
        for (int k = 0; k < 3; ++k) {
            idx += 1;
            if (idx == 3) idx = 0;
            const double fj = idx == 0 ? f0 - 3.0 : (idx == 1 ? -x[0] + x[1] + 1.0 : gamma - f0);
            if (fj > 0.0) {
                g[0] = idx == 0 ? 1.0 : -1.0;
                g[1] = idx == 2 ? -1.0 : 1.0;
                cut = single_cut(fj);
                return false;
            }
        }
        gamma = f0;
        g[0] = -1.0; g[1] = -1.0; cut = single_cut(0.0);
        return true;
    }
};

__device__ inline void example4_grad(int idx, double sign3, double (&g)[2]) {
    switch (idx) {
        case 0: g[0] = -1.0; g[1] = 0.0; break;
        case 1: g[0] = 0.0; g[1] = -1.0; break;
        case 2: g[0] = 1.0; g[1] = 1.0; break;
        default: g[0] = -2.0 * sign3; g[1] = 3.0 * sign3; break;
    }
}

struct OrExample4 {  // src/example4.rs:10-57
    static constexpr int kN = 2;
    int idx;
    __device__ void init(const BatchArgs&, long long) { idx = -1; }
    __device__ bool assess_optim(const double (&x)[2], double& gamma, double (&g)[2], Cut& cut) {
        const double f0 = 2.0 * x[0] - 3.0 * x[1];
        for (int k = 0; k < 4; ++k) {
            idx += 1;
            if (idx == 4) idx = 0;
            double fj;
            switch (idx) {
                case 0: fj = -x[0] - 1.0; break;
                case 1: fj = -x[1] - 2.0; break;
                case 2: fj = x[0] + x[1] - 1.0; break;
                default: fj = gamma - f0; break;
            }
            if (fj > 0.0) {
                example4_grad(idx, 1.0, g);
                cut = single_cut(fj);
                return false;
            }
        }
        gamma = f0;
        g[0] = -2.0; g[1] = 3.0; cut = single_cut(0.0);
        return true;
    }
};

struct OrExample3 {  // src/example3.rs:10-60 (feasibility; `target` is what OracleFeas::update sets)
    static constexpr int kN = 2;
    int idx;
    double target;
    __device__ void init(const BatchArgs& a, long long pid) {
        idx = -1;
        target = a.params ? a.params[pid * a.params_stride] : -1e100;
    }
    __device__ void update(double gamma) { target = gamma; }   // OracleFeas::update, src/example3.rs:56-58
    __device__ bool assess_feas(const double (&x)[2], double (&g)[2], Cut& cut) {
        for (int k = 0; k < 4; ++k) {
            idx += 1;
            if (idx == 4) idx = 0;
            double fj;
            switch (idx) {
                case 0: fj = -x[0] - 1.0; break;
                case 1: fj = -x[1] - 2.0; break;
                case 2: fj = x[0] + x[1] - 1.0; break;
                default: fj = 2.0 * x[0] - 3.0 * x[1] - target; break;
            }
            if (fj > 0.0) {
                example4_grad(idx, -1.0, g);  // constraint 3 here has gradient (2, -3)
                cut = single_cut(fj);
                return true;
            }
        }
        return false;
    }
};

struct OrQuasicvx {  // src/quasicvx.rs:11-51
    static constexpr int kN = 2;
    int idx;
    __device__ void init(const BatchArgs&, long long) { idx = -1; }
    __device__ bool assess_optim(const double (&x)[2], double& gamma, double (&g)[2], Cut& cut) {
        const double sqrtx = x[0], logy = x[1];
        for (int k = 0; k < 2; ++k) {
            idx += 1;
            if (idx == 2) idx = 0;
            const double fv = idx == 0 ? sqrtx * sqrtx - logy : -sqrtx + gamma * glibc::exp(logy);
            if (fv > 0.0) {
                if (idx == 0) { g[0] = 2.0 * sqrtx; g[1] = -1.0; }
                else { g[0] = -1.0; g[1] = gamma * glibc::exp(logy); }
                cut = single_cut(fv);
                return false;
            }
        }
        gamma = sqrtx / glibc::exp(logy);
        g[0] = -1.0; g[1] = sqrtx; cut = single_cut(0.0);
        return true;
    }
};

struct OrProfit {  // src/oracles/profit_oracle.rs:42-114
    static constexpr int kN = 2;
    int idx;
    double log_p_scale, log_k, v0, v1, a0, a1, log_cobb, vx, q0, q1;
    __device__ void setup(double p, double scale, double limit, double a0_, double a1_, double v0_, double v1_) {  // :54-68
        idx = -1;
        log_p_scale = glibc::log(p * scale);
        log_k = glibc::log(limit);
        a0 = a0_; a1 = a1_; v0 = v0_; v1 = v1_;
        log_cobb = 0.0; vx = 0.0; q0 = 0.0; q1 = 0.0;
    }
    __device__ void init(const BatchArgs& a, long long pid) {
        const double* p = a.params + pid * a.params_stride;
        setup(p[0], p[1], p[2], p[3], p[4], p[5], p[6]);
    }
    __device__ bool feas(const double (&y)[2], double gamma, double (&g)[2], double& fj_out) {  // :70-101
        for (int k = 0; k < 2; ++k) {
            idx += 1;
            if (idx == 2) idx = 0;
            double fj;
            if (idx == 0) {
                fj = y[0] - log_k;
            } else {
                double dotv = 0.0;  // elasticities.dot(y): sequential from 0.0 (src/arr.rs:443-451)
                dotv += a0 * y[0];
                dotv += a1 * y[1];
                log_cobb = log_p_scale + dotv;
                q0 = v0 * glibc::exp(y[0]);
                q1 = v1 * glibc::exp(y[1]);
                vx = q0 + q1;
                fj = glibc::log(gamma + vx) - log_cobb;
            }
            if (fj > 0.0) {
                if (idx == 0) { g[0] = 1.0; g[1] = 0.0; }
                else {
                    const double inv = 1.0 / (gamma + vx);  // Arr / f64 multiplies by the reciprocal (src/arr.rs:241-249)
                    g[0] = q0 * inv - a0;
                    g[1] = q1 * inv - a1;
                }
                fj_out = fj;
                return true;
            }
        }
        return false;
    }
    __device__ bool assess_optim(const double (&y)[2], double& gamma, double (&g)[2], Cut& cut) {  // :104-113
        double fj;
        if (feas(y, gamma, g, fj)) { cut = single_cut(fj); return false; }
        const double exp_val = glibc::exp(log_cobb);
        gamma = exp_val - vx;
        const double inv = 1.0 / exp_val;
        g[0] = q0 * inv - a0;
        g[1] = q1 * inv - a1;
        cut = single_cut(0.0);
        return true;
    }
};

struct OrProfitRb {  // src/oracles/profit_oracle.rs:116-161
    static constexpr int kN = 2;
    OrProfit omega;
    double uie0, uie1, a0, a1;
    __device__ void init(const BatchArgs& a, long long pid) {
        const double* p = a.params + pid * a.params_stride;
        const double e1 = p[7], e2 = p[8], e3 = p[9], e4 = p[10], e5 = p[11];
        uie0 = e1; uie1 = e2;
        a0 = p[3]; a1 = p[4];
        omega.setup(p[0] - e3, p[1], p[2] - e4, p[3], p[4], p[5] + e5, p[6] + e5);
    }
    __device__ bool assess_optim(const double (&y)[2], double& gamma, double (&g)[2], Cut& cut) {
        double r0 = a0, r1 = a1;
        r0 += (y[0] > 0.0) ? -uie0 : uie0;
        r1 += (y[1] > 0.0) ? -uie1 : uie1;
        omega.a0 = r0; omega.a1 = r1;
        return omega.assess_optim(y, gamma, g, cut);
    }
};

struct OrProfitQ {  // src/oracles/profit_oracle.rs:163-198
    static constexpr int kN = 2;
    OrProfit omega;
    double yd[2];
    __device__ void init(const BatchArgs& a, long long pid) { omega.init(a, pid); yd[0] = 0.0; yd[1] = 0.0; }
    __device__ bool assess_optim_q(const double (&y)[2], double& gamma, bool retry, double (&g)[2], Cut& cut, double (&xq)[2],
                                   bool& more_alt) {
        if (!retry) {
            double fj;
            if (omega.feas(y, gamma, g, fj)) {
                cut = single_cut(fj);
                xq[0] = y[0]; xq[1] = y[1];
                more_alt = true;
                return false;
            }
            double d0 = round(glibc::exp(y[0])), d1 = round(glibc::exp(y[1]));  // f64::round: half away from zero == CUDA round()
            if (d0 == 0.0) d0 = 1.0;
            if (d1 == 0.0) d1 = 1.0;
            yd[0] = glibc::log(d0); yd[1] = glibc::log(d1);
        }
        const bool shrunk = omega.assess_optim(yd, gamma, g, cut);
        double dotv = 0.0;  // grad.dot(&(&yd - y))
        dotv += g[0] * (yd[0] - y[0]);
        dotv += g[1] * (yd[1] - y[1]);
        cut = single_cut(cut.b0 + dotv);
        xq[0] = yd[0]; xq[1] = yd[1];
        more_alt = !retry;
        return shrunk;
    }
};

struct OrExample2 {  // tests/example2_tests.rs:12-48
    static constexpr int kN = 2;
    int idx;
    __device__ void init(const BatchArgs&, long long) { idx = -1; }
    __device__ bool assess_feas(const double (&x)[2], double (&g)[2], Cut& cut) {
        for (int k = 0; k < 2; ++k) {
            idx += 1;
            if (idx == 2) idx = 0;
            const double fj = idx == 0 ? x[0] + x[1] - 3.0 : -x[0] + x[1] + 1.0;
            if (fj > 0.0) {
                g[0] = idx == 0 ? 1.0 : -1.0;
                g[1] = 1.0;
                cut = single_cut(fj);
                return true;
            }
        }
        return false;
    }
};

struct OrHalfPlane {  // tests/cutting_plane_tests.rs:12-28
    static constexpr int kN = 2;
    __device__ void init(const BatchArgs&, long long) {}
    __device__ bool assess_feas(const double (&x)[2], double (&g)[2], Cut& cut) {
        const double fj = x[0] + x[1] - 3.0;
        if (fj > 0.0) { g[0] = 1.0; g[1] = 1.0; cut = single_cut(fj); return true; }
        return false;
    }
};

struct OrAlwaysCut {  // tests/cutting_plane_tests.rs:33-42, 91-99, 104-122
    static constexpr int kN = 2;
    __device__ void init(const BatchArgs&, long long) {}
    __device__ bool assess_feas(const double (&)[2], double (&g)[2], Cut& cut) {
        g[0] = 1.0; g[1] = 1.0; cut = single_cut(1.0);
        return true;
    }
    __device__ bool assess_optim(const double (&)[2], double&, double (&g)[2], Cut& cut) {
        g[0] = 1.0; g[1] = 1.0; cut = single_cut(1.0);
        return false;
    }
    __device__ bool assess_optim_q(const double (&x)[2], double&, bool, double (&g)[2], Cut& cut, double (&xq)[2], bool& more_alt) {
        g[0] = 1.0; g[1] = 1.0; cut = single_cut(1.0);
        xq[0] = x[0]; xq[1] = x[1];
        more_alt = true;
        return false;
    }
};

struct OrOptimBox {  // tests/cutting_plane_tests.rs:47-73 (optim) and :197-264 (optim_q)
    static constexpr int kN = 2;
    __device__ void init(const BatchArgs&, long long) {}
    __device__ bool assess_optim(const double (&x)[2], double& gamma, double (&g)[2], Cut& cut) {
        const double f0 = x[0] + x[1];
        const double f1 = x[0] - 1.0;
        if (f1 > 0.0) { g[0] = 1.0; g[1] = 0.0; cut = single_cut(f1); return false; }
        const double f2 = x[1] - 1.0;
        if (f2 > 0.0) { g[0] = 0.0; g[1] = 1.0; cut = single_cut(f2); return false; }
        const double f3 = f0 - gamma;
        g[0] = -1.0; g[1] = -1.0;
        if (f3 < 0.0) { cut = single_cut(-f3); return false; }
        cut = single_cut(0.0);
        return true;
    }
    __device__ bool assess_optim_q(const double (&x)[2], double& gamma, bool retry, double (&g)[2], Cut& cut, double (&xq)[2],
                                   bool& more_alt) {
        xq[0] = x[0]; xq[1] = x[1];
        more_alt = true;
        const double f0 = x[0] + x[1];
        const double f1 = x[0] - 1.0;
        if (f1 > 0.0) { g[0] = 1.0; g[1] = 0.0; cut = single_cut(f1); return false; }
        const double f2 = x[1] - 1.0;
        if (f2 > 0.0) { g[0] = 0.0; g[1] = 1.0; cut = single_cut(f2); return false; }
        const double f3 = f0 - gamma;
        if (f3 < 0.0) { g[0] = -1.0; g[1] = -1.0; cut = single_cut(-f3); return false; }
        xq[0] = round(x[0]); xq[1] = round(x[1]);
        more_alt = !retry;
        const double f1q = xq[0] - 1.0;
        if (f1q > 0.0) { g[0] = 1.0; g[1] = 0.0; cut = single_cut(f1q); return false; }
        const double f2q = xq[1] - 1.0;
        if (f2q > 0.0) { g[0] = 0.0; g[1] = 1.0; cut = single_cut(f2q); return false; }
        const double f3q = xq[0] + xq[1] - gamma;
        g[0] = -1.0; g[1] = -1.0;
        if (f3q < 0.0) { cut = single_cut(-f3q); return false; }
        gamma = xq[0] + xq[1];
        cut = single_cut(0.0);
        return true;
    }
};

// tests/lmi_tests.rs:121-172 — MyLmiOracle: minimise c·x subject to L LMIs B_b - sum_k F_bk x_k >= 0, round-robin over the
// L blocks and the objective cut.  One thread per problem: n = 3 variables, up to 3 blocks of size m <= 4; every block is
// LMIOracle::assess_feas (src/oracles/lmi_oracle.rs:24-44) over LDLTMgr (src/oracles/ldlt_mgr.rs:29-55, 98-124), all in
// local arrays.  table = c[3], then per block F_0, F_1, F_2 (m×m each), B (m×m); aux[0] = L, aux[1..L] = block sizes.
struct OrLmiLinOpt {
    static constexpr int kN = 3;
    static constexpr int kMaxM = 4, kMaxBlocks = 3;
    int idx, nblocks;
    int msize[kMaxBlocks];
    const double* fbase[kMaxBlocks];
    const double* c;
    __device__ void init(const BatchArgs& a, long long) {
        idx = -1;
        nblocks = a.aux[0];
        c = a.table;
        const double* t = a.table + kN;
        for (int b = 0; b < kMaxBlocks; ++b) {
            msize[b] = b < nblocks ? a.aux[1 + b] : 0;
            fbase[b] = t;
            t += (kN + 1) * msize[b] * msize[b];
        }
    }
    // LMIOracle::assess_feas for block b: true if a cut (g, ep) was produced
    __device__ bool block_feas(int b, const double (&x)[3], double (&g)[3], double& ep) const {
        const int m = msize[b];
        const double* F = fbase[b];
        const double* B = F + kN * m * m;
        double T[kMaxM * kMaxM];
        double wit[kMaxM];
        auto elem = [&](int i, int j) {   // lmi_oracle.rs:28-34
            double s = B[i * m + j];
            for (int k = 0; k < kN; ++k) s -= F[k * m * m + i * m + j] * x[k];
            return s;
        };
        int pos = 0;
        for (int i = 0; i < m; ++i) {      // ldlt_mgr.rs:29-55
            double diag = elem(i, 0);
            for (int j = 0; j < i; ++j) {
                T[j * kMaxM + i] = diag;
                T[i * kMaxM + j] = diag / T[j * kMaxM + j];
                double sum = 0.0;
                for (int k = 0; k <= j; ++k) sum += T[i * kMaxM + k] * T[k * kMaxM + (j + 1)];
                diag = elem(i, j + 1) - sum;
            }
            T[i * kMaxM + i] = diag;
            if (diag <= 0.0) { pos = i + 1; break; }
        }
        if (pos == 0) return false;
        const int mm = pos - 1;             // witness, ldlt_mgr.rs:98-111
        wit[mm] = 1.0;
        for (int i = mm; i >= 1; --i) {
            double sum = 0.0;
            for (int k = i; k < pos; ++k) sum += T[k * kMaxM + (i - 1)] * wit[k];
            wit[i - 1] = -sum;
        }
        ep = -T[mm * kMaxM + mm];
        for (int k = 0; k < kN; ++k) {      // sym_quad, ldlt_mgr.rs:115-124
            double r = 0.0;
            for (int i = 0; i < pos; ++i)
                for (int j = 0; j < pos; ++j) r += wit[i] * F[k * m * m + i * m + j] * wit[j];
            g[k] = r;
        }
        return true;
    }
    __device__ bool assess_optim(const double (&x)[3], double& gamma, double (&g)[3], Cut& cut) {
        double f0 = 0.0;                    // self.c.dot(xc)
        for (int k = 0; k < kN; ++k) f0 += c[k] * x[k];
        const int nc = nblocks + 1;
        for (int t = 0; t < nc; ++t) {
            idx = (idx == nc - 1) ? 0 : idx + 1;
            if (idx < nblocks) {
                double ep;
                if (block_feas(idx, x, g, ep)) { cut = single_cut(ep); return false; }
            } else {
                const double fj = f0 - gamma;
                if (fj > 0.0) {
                    for (int k = 0; k < kN; ++k) g[k] = c[k];
                    cut = single_cut(fj);
                    return false;
                }
                gamma = f0;
            }
        }
        for (int k = 0; k < kN; ++k) g[k] = c[k];
        cut = single_cut(0.0);
        return true;
    }
};

// ------------------------------------------------------------------------------------------------ kernel
template <int N>
__device__ __forceinline__ void copy_vec(double (&dst)[N], const double (&src)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) dst[i] = src[i];
}

template <int N, class Space, class Oracle, int MODE>
__global__ void __launch_bounds__(128) batch_small_kernel(const BatchArgs a) {
    const CalcConsts cc = CalcConsts::make(N, a.use_parallel_cut);
    const int lane = threadIdx.x & 31;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.queue, 32ULL);
        base = __shfl_sync(0xffffffffu, base, 0);
        if ((long long)base >= a.count) break;
        const long long pid = (long long)base + lane;
        if (pid < a.count) {
            Space sp;
            sp.kappa = a.kappa0[pid * a.kappa0_stride];
            sp.tsq = 0.0;
            sp.no_defer = a.no_defer_trick != 0;
#pragma unroll
            for (int i = 0; i < N; ++i) sp.xc[i] = a.xc0[pid * a.xc0_stride + i];
#pragma unroll
            for (int i = 0; i < N; ++i)
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    double v = (i == j) ? 1.0 : 0.0;
                    if (a.mq0) v = a.mq0[pid * a.mq0_stride + i * N + j];
                    else if (a.diag0 && i == j) v = a.diag0[pid * a.diag0_stride + i];
                    sp.m[i * N + j] = v;
                }
            double gamma = (MODE != ELL_MODE_FEAS && a.gamma0) ? a.gamma0[pid * a.gamma0_stride] : 0.0;
            Oracle om;
            om.init(a, pid);

            bool has_best = false;
            double best[N];
#pragma unroll
            for (int i = 0; i < N; ++i) best[i] = 0.0;
            long long niter = 0;
            int exit_status = kExitMaxIt;
            double g[N];
            Cut cut;
            if constexpr (MODE == ELL_MODE_OPTIM) {  // src/cutting_plane.rs:286-313
                for (; niter < a.max_iters; ++niter) {
                    const bool shrunk = om.assess_optim(sp.xc, gamma, g, cut);
                    int status;
                    if (shrunk) {
                        has_best = true;
                        copy_vec<N>(best, sp.xc);
                        status = sp.update(cc, kCentral, g, cut);
                    } else {
                        status = sp.update(cc, kBias, g, cut);
                    }
                    if (status != kSuccess) { exit_status = status; break; }
                    if (sp.tsq < a.tolerance) { exit_status = kExitTol; break; }
                }
            } else if constexpr (MODE == ELL_MODE_FEAS) {  // :205-227
                for (; niter < a.max_iters; ++niter) {
                    if (!om.assess_feas(sp.xc, g, cut)) {
                        has_best = true;
                        copy_vec<N>(best, sp.xc);
                        exit_status = kExitFeasible;
                        break;
                    }
                    const int status = sp.update(cc, kBias, g, cut);
                    if (status != kSuccess) { exit_status = status; break; }
                    if (sp.tsq < a.tolerance) { exit_status = kExitTol; break; }
                }
            } else if constexpr (MODE == ELL_MODE_BSEARCH) {  // bsearch :441-466 over BSearchAdaptor::assess_bs :409-418
                const double* p = a.params + pid * a.params_stride;
                double lower = gamma, upper = p[0];
                const double u_orig = upper, outer_tol = p[1];
                const long long outer_max = (long long)p[2];
                om.target = -1e100;   // MyOracle3::default(); params[] belong to the search here, not to the oracle
                exit_status = kExitMaxIt;
                for (niter = 0; niter < outer_max; ++niter) {
                    const double tau = (upper - lower) / 2.0;
                    if (tau < outer_tol) { exit_status = kExitTol; break; }
                    double gm = lower;
                    gm += tau;
                    Space trial = sp;          // let mut space = self.space.clone();
                    om.update(gm);
                    bool feasible = false;
                    for (long long it = 0; it < a.max_iters; ++it) {   // cutting_plane_feas :205-227
                        if (!om.assess_feas(trial.xc, g, cut)) { feasible = true; break; }
                        const int status = trial.update(cc, kBias, g, cut);
                        if (status != kSuccess || trial.tsq < a.tolerance) break;
                    }
                    if (feasible) {
                        copy_vec<N>(sp.xc, trial.xc);   // self.space.set_xc(x)
                        upper = gm;
                    } else {
                        lower = gm;
                    }
                }
                has_best = upper != u_orig;
                copy_vec<N>(best, sp.xc);
                gamma = upper;
            } else {  // optim_q :331-374
                bool retry = false;
                for (; niter < a.max_iters; ++niter) {
                    double xq[N];
                    bool more_alt;
                    const bool shrunk = om.assess_optim_q(sp.xc, gamma, retry, g, cut, xq, more_alt);
                    if (shrunk) {
                        has_best = true;
                        copy_vec<N>(best, xq);
                        retry = false;
                    }
                    const int status = sp.update(cc, kQ, g, cut);
                    if (status == kSuccess) {
                        retry = false;
                    } else if (status == kNoSoln) {
                        exit_status = kExitNoSoln;
                        break;
                    } else if (status == kNoEffect) {
                        if (!more_alt) { exit_status = kExitNoEffect; break; }
                        retry = true;
                    }
                    if (sp.tsq < a.tolerance) { exit_status = kExitTol; break; }
                }
            }
            a.num_iters[pid] = niter;
            a.exit_status[pid] = exit_status;
            if (a.has_xbest) a.has_xbest[pid] = has_best ? 1 : 0;
            if (a.xbest && has_best) {
#pragma unroll
                for (int i = 0; i < N; ++i) a.xbest[pid * N + i] = best[i];
            }
            if (a.gamma) a.gamma[pid] = gamma;
            if (a.tsq) a.tsq[pid] = sp.tsq;
            if (a.kappa) a.kappa[pid] = sp.kappa;
            if (a.xc) {
#pragma unroll
                for (int i = 0; i < N; ++i) a.xc[pid * N + i] = sp.xc[i];
            }
        }
    }
}

#ifdef __CUDACC__   // the launch itself: tests/cpp/emu_batch_small.cpp supplies a host emulation of launch_kind under the same dispatch table
template <class Oracle, int MODE>
int launch_kind(const BatchArgs& a, cudaStream_t s, int sm_count) {
    const long long warps_needed = (a.count + 31) / 32;
    long long blocks = (warps_needed + 3) / 4;
    const long long cap = (long long)sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    constexpr int N = Oracle::kN;
    if (a.kind == ELL_KIND_ELLSTABLE)
        batch_small_kernel<N, EllStableReg<N>, Oracle, MODE><<<(unsigned)blocks, 128, 0, s>>>(a);
    else
        batch_small_kernel<N, EllReg<N>, Oracle, MODE><<<(unsigned)blocks, 128, 0, s>>>(a);
    ELL_CUDA_TRY(cudaGetLastError());
    return ELLCUDA_OK;
}

#endif  // __CUDACC__

}  // namespace

#ifdef __CUDACC__
__global__ void math_probe_kernel(int which, const double* __restrict__ x, double* __restrict__ y, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = which == 0 ? glibc::exp(x[i]) : glibc::log(x[i]);
}

int launch_math_probe(int which, const double* x, double* y, long long n, cudaStream_t s) {
    math_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(which, x, y, n);
    ELL_CUDA_TRY(cudaGetLastError());
    return ELLCUDA_OK;
}

#endif  // __CUDACC__

int launch_batch_small(const BatchArgs& a, int oracle_id, cudaStream_t s, int sm_count, long long* launches) {
    const int m = a.mode;
    if (oracle_id == ELL_ORACLE_LMI_LINOPT) {
        if (a.n != 3 || m != ELL_MODE_OPTIM || !a.table || a.aux[0] < 1 || a.aux[0] > OrLmiLinOpt::kMaxBlocks) {
            set_error("LMI_LINOPT on the device: n = 3, cutting_plane_optim, 1..3 blocks of size <= 4");
            return ELLCUDA_E_ARG;
        }
        long long need = 3;
        for (int b = 0; b < a.aux[0]; ++b) {
            if (a.aux[1 + b] < 1 || a.aux[1 + b] > OrLmiLinOpt::kMaxM) { set_error("LMI_LINOPT: block size must be 1..4"); return ELLCUDA_E_ARG; }
            need += 4LL * a.aux[1 + b] * a.aux[1 + b];
        }
        if (a.table_len < need) { set_error("LMI_LINOPT: table too short"); return ELLCUDA_E_ARG; }
        const int rc = launch_kind<OrLmiLinOpt, ELL_MODE_OPTIM>(a, s, sm_count);
        if (rc == ELLCUDA_OK && launches) *launches += 1;
        return rc;
    }
    if (a.n != 2) { set_error("this oracle is defined for n = 2"); return ELLCUDA_E_ARG; }
    int rc = ELLCUDA_E_ARG;
#define ELL_CASE(ID, OR, MODE) \
    if (oracle_id == (ID) && m == (MODE)) { rc = launch_kind<OR, MODE>(a, s, sm_count); goto done; }
    ELL_CASE(ELL_ORACLE_EXAMPLE1, OrExample1, ELL_MODE_OPTIM)
    ELL_CASE(ELL_ORACLE_EXAMPLE1_RR, OrExample1RR, ELL_MODE_OPTIM)
    ELL_CASE(ELL_ORACLE_EXAMPLE4, OrExample4, ELL_MODE_OPTIM)
    ELL_CASE(ELL_ORACLE_QUASICVX, OrQuasicvx, ELL_MODE_OPTIM)
    ELL_CASE(ELL_ORACLE_PROFIT, OrProfit, ELL_MODE_OPTIM)
    ELL_CASE(ELL_ORACLE_PROFIT_RB, OrProfitRb, ELL_MODE_OPTIM)
    ELL_CASE(ELL_ORACLE_PROFIT_Q, OrProfitQ, ELL_MODE_OPTIM_Q)
    ELL_CASE(ELL_ORACLE_EXAMPLE2, OrExample2, ELL_MODE_FEAS)
    ELL_CASE(ELL_ORACLE_EXAMPLE3, OrExample3, ELL_MODE_FEAS)
    ELL_CASE(ELL_ORACLE_EXAMPLE3, OrExample3, ELL_MODE_BSEARCH)
    ELL_CASE(ELL_ORACLE_HALFPLANE, OrHalfPlane, ELL_MODE_FEAS)
    ELL_CASE(ELL_ORACLE_ALWAYS_CUT, OrAlwaysCut, ELL_MODE_FEAS)
    ELL_CASE(ELL_ORACLE_ALWAYS_CUT, OrAlwaysCut, ELL_MODE_OPTIM)
    ELL_CASE(ELL_ORACLE_ALWAYS_CUT, OrAlwaysCut, ELL_MODE_OPTIM_Q)
    ELL_CASE(ELL_ORACLE_OPTIM_BOX, OrOptimBox, ELL_MODE_OPTIM)
    ELL_CASE(ELL_ORACLE_OPTIM_BOX, OrOptimBox, ELL_MODE_OPTIM_Q)
#undef ELL_CASE
    set_error("oracle / mode combination not available in the n=2 kernel");
    return ELLCUDA_E_ARG;
done:
    if (rc == ELLCUDA_OK && launches) *launches += 1;
    return rc;
}

}  // namespace ellcuda
