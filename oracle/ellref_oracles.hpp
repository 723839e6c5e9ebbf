// ellref_oracles.hpp — CPU restatement of the reference's oracles and test fixtures.
// TEST INFRASTRUCTURE, NOT PRODUCT (see ellref.hpp).  Each struct cites the lines it follows.
#pragma once
#include "ellref.hpp"

namespace ellref {

using CutPair = std::pair<Vec, Cut>;
using OptimAnswer = std::pair<CutPair, bool>;
using FeasAnswer = std::optional<CutPair>;

// ---- src/example1.rs:8-30 ------------------------------------------------------------------------
struct Example1 {
    OptimAnswer assess_optim(const Vec& xc, double& gamma) {
        const double x = xc[0], y = xc[1];
        const double f0 = x + y;
        const double f1 = f0 - 3.0;
        if (f1 > 0.0) return {{{1.0, 1.0}, Cut::single(f1)}, false};
        const double f2 = -x + y + 1.0;
        if (f2 > 0.0) return {{{-1.0, 1.0}, Cut::single(f2)}, false};
        const double f3 = gamma - f0;
        if (f3 > 0.0) return {{{-1.0, -1.0}, Cut::single(f3)}, false};
        gamma = f0;
        return {{{-1.0, -1.0}, Cut::single(0.0)}, true};
    }
};

// ---- src/example1_rr.rs:13-53 --------------------------------------------------------------------
struct Example1RR {
    int idx = -1;
    OptimAnswer assess_optim(const Vec& xc, double& gamma) {
        const double x = xc[0], y = xc[1];
        const double f0 = x + y;
        for (int k = 0; k < 3; ++k) {
            idx += 1;
            if (idx == 3) idx = 0;
            double fj;
            switch (idx) {
                case 0: fj = f0 - 3.0; break;
                case 1: fj = -x + y + 1.0; break;
                default: fj = gamma - f0; break;
            }
            if (fj > 0.0) {
                Vec g = idx == 0 ? Vec{1.0, 1.0} : (idx == 1 ? Vec{-1.0, 1.0} : Vec{-1.0, -1.0});
                return {{g, Cut::single(fj)}, false};
            }
        }
        gamma = f0;
        return {{{-1.0, -1.0}, Cut::single(0.0)}, true};
    }
};

// ---- src/example4.rs:10-57 -----------------------------------------------------------------------
struct Example4 {
    int idx = -1;
    OptimAnswer assess_optim(const Vec& xc, double& gamma) {
        const double x = xc[0], y = xc[1];
        const double f0 = 2.0 * x - 3.0 * y;
        for (int k = 0; k < 4; ++k) {
            idx += 1;
            if (idx == 4) idx = 0;
            double fj;
            switch (idx) {
                case 0: fj = -x - 1.0; break;
                case 1: fj = -y - 2.0; break;
                case 2: fj = x + y - 1.0; break;
                default: fj = gamma - f0; break;
            }
            if (fj > 0.0) {
                Vec g;
                switch (idx) {
                    case 0: g = {-1.0, 0.0}; break;
                    case 1: g = {0.0, -1.0}; break;
                    case 2: g = {1.0, 1.0}; break;
                    default: g = {-2.0, 3.0}; break;
                }
                return {{g, Cut::single(fj)}, false};
            }
        }
        gamma = f0;
        return {{{-2.0, 3.0}, Cut::single(0.0)}, true};
    }
};

// ---- src/example3.rs:10-60 (feasibility, used under bsearch) --------------------------------------
struct Example3 {
    int idx = -1;
    double target = -1e100;
    FeasAnswer assess_feas(const Vec& xc) {
        const double x = xc[0], y = xc[1];
        for (int k = 0; k < 4; ++k) {
            idx += 1;
            if (idx == 4) idx = 0;
            double fj;
            switch (idx) {
                case 0: fj = -x - 1.0; break;
                case 1: fj = -y - 2.0; break;
                case 2: fj = x + y - 1.0; break;
                default: fj = 2.0 * x - 3.0 * y - target; break;
            }
            if (fj > 0.0) {
                Vec g;
                switch (idx) {
                    case 0: g = {-1.0, 0.0}; break;
                    case 1: g = {0.0, -1.0}; break;
                    case 2: g = {1.0, 1.0}; break;
                    default: g = {2.0, -3.0}; break;
                }
                return CutPair{g, Cut::single(fj)};
            }
        }
        return std::nullopt;
    }
    void update(double gamma) { target = gamma; }
};

// ---- src/quasicvx.rs:11-51 -----------------------------------------------------------------------
struct Quasicvx {
    int idx = -1;
    OptimAnswer assess_optim(const Vec& xc, double& gamma) {
        const double sqrtx = xc[0], logy = xc[1];
        for (int k = 0; k < 2; ++k) {
            idx += 1;
            if (idx == 2) idx = 0;
            const double fv = idx == 0 ? sqrtx * sqrtx - logy : -sqrtx + gamma * std::exp(logy);
            if (fv > 0.0) {
                Vec g = idx == 0 ? Vec{2.0 * sqrtx, -1.0} : Vec{-1.0, gamma * std::exp(logy)};
                return {{g, Cut::single(fv)}, false};
            }
        }
        gamma = sqrtx / std::exp(logy);
        return {{{-1.0, sqrtx}, Cut::single(0.0)}, true};
    }
};

// ---- src/oracles/profit_oracle.rs:42-114 ----------------------------------------------------------
struct Profit {
    int idx = -1;
    double log_p_scale, log_k;
    double v[2];  // price_out
    double a[2];  // elasticities
    double log_cobb = 0.0, vx = 0.0;
    double q[2] = {0.0, 0.0};
    Profit(double p, double scale, double limit, const double* a_, const double* v_) {  // :54-68
        log_p_scale = std::log(p * scale);
        log_k = std::log(limit);
        a[0] = a_[0]; a[1] = a_[1];
        v[0] = v_[0]; v[1] = v_[1];
    }
    std::optional<std::pair<Vec, double>> feas(const Vec& y, double& gamma) {  // :70-101
        for (int k = 0; k < 2; ++k) {
            idx += 1;
            if (idx == 2) idx = 0;
            double fj;
            if (idx == 0) {
                fj = y[0] - log_k;
            } else {
                log_cobb = log_p_scale + dot(a, y.data(), 2);
                q[0] = v[0] * std::exp(y[0]);
                q[1] = v[1] * std::exp(y[1]);
                vx = q[0] + q[1];
                fj = std::log(gamma + vx) - log_cobb;
            }
            if (fj > 0.0) {
                if (idx == 0) return std::pair<Vec, double>{Vec{1.0, 0.0}, fj};
                const double inv = 1.0 / (gamma + vx);  // Arr / f64 multiplies by the reciprocal, arr.rs:241-249
                return std::pair<Vec, double>{Vec{q[0] * inv - a[0], q[1] * inv - a[1]}, fj};
            }
        }
        return std::nullopt;
    }
    OptimAnswer assess_optim(const Vec& y, double& gamma) {  // :104-113
        if (auto c = feas(y, gamma)) return {{c->first, Cut::single(c->second)}, false};
        const double exp_val = std::exp(log_cobb);
        gamma = exp_val - vx;
        const double inv = 1.0 / exp_val;
        return {{{q[0] * inv - a[0], q[1] * inv - a[1]}, Cut::single(0.0)}, true};
    }
};

// ---- src/oracles/profit_oracle.rs:116-161 ---------------------------------------------------------
struct ProfitRb {
    double uie[2];
    Profit omega;
    double a[2];
    static Profit make(double p, double scale, double limit, const double* a_, const double* v_, const double* e) {
        const double v2[2] = {v_[0] + e[4], v_[1] + e[4]};
        return Profit(p - e[2], scale, limit - e[3], a_, v2);
    }
    ProfitRb(double p, double scale, double limit, const double* a_, const double* v_, const double* e)
        : omega(make(p, scale, limit, a_, v_, e)) {
        uie[0] = e[0]; uie[1] = e[1];
        a[0] = a_[0]; a[1] = a_[1];
    }
    OptimAnswer assess_optim(const Vec& y, double& gamma) {  // :146-160
        double a_rb[2] = {a[0], a[1]};
        for (int i = 0; i < 2; ++i) a_rb[i] += (y[i] > 0.0) ? -uie[i] : uie[i];
        omega.a[0] = a_rb[0]; omega.a[1] = a_rb[1];
        return omega.assess_optim(y, gamma);
    }
};

// ---- src/oracles/profit_oracle.rs:163-198 ---------------------------------------------------------
struct ProfitQ {
    Profit omega;
    Vec yd{0.0, 0.0};
    ProfitQ(double p, double scale, double limit, const double* a_, const double* v_) : omega(p, scale, limit, a_, v_) {}
    QAnswer assess_optim_q(const Vec& y, double& gamma, bool retry) {
        if (!retry) {
            if (auto c = omega.feas(y, gamma)) return {{c->first, Cut::single(c->second)}, false, y, true};
            Vec x_disc{std::round(std::exp(y[0])), std::round(std::exp(y[1]))};  // f64::round = half away from zero
            if (x_disc[0] == 0.0) x_disc[0] = 1.0;
            if (x_disc[1] == 0.0) x_disc[1] = 1.0;
            yd = {std::log(x_disc[0]), std::log(x_disc[1])};
        }
        OptimAnswer o = omega.assess_optim(yd, gamma);
        const Vec& grad = o.first.first;
        const Vec d{yd[0] - y[0], yd[1] - y[1]};
        const double beta = o.first.second.beta0 + dot(grad, d);
        return {{grad, Cut::single(beta)}, o.second, yd, !retry};
    }
};

// ---- src/oracles/lowpass_oracle.rs:7-150 ----------------------------------------------------------
inline Vec lowpass_spectrum(std::size_t ndim) {  // :24-33, flattened mdim×ndim row-major
    const std::size_t mdim = 15 * ndim;
    const Vec omega = linspace(0.0, 3.14159265358979323846264338327950288, mdim);
    Vec s(mdim * ndim, 0.0);
    for (std::size_t i = 0; i < mdim; ++i) {
        s[i * ndim] = 1.0;
        for (std::size_t j = 1; j < ndim; ++j) s[i * ndim + j] = 2.0 * std::cos(omega[i] * static_cast<double>(j));
    }
    return s;
}
inline int lowpass_band_edge(std::size_t ndim, double w) {  // :35-36
    const std::size_t mdim = 15 * ndim;
    return static_cast<int>(std::floor(w * static_cast<double>(mdim - 1))) + 1;
}

struct Lowpass {
    bool more_alt = true;
    int idx1;
    const double* spectrum;  // mdim × ndim
    std::size_t mdim, ndim;
    int nwpass, nwstop;
    double lp_sq, up_sq, sp_sq;
    int idx2, idx3;
    double fmax = -std::numeric_limits<double>::infinity();
    int kmax = -1;
    Lowpass(std::size_t ndim_, const double* table, double wpass, double wstop, double lp, double up, double sp)
        : idx1(-1), spectrum(table), mdim(15 * ndim_), ndim(ndim_), lp_sq(lp), up_sq(up), sp_sq(sp) {  // :23-52
        nwpass = lowpass_band_edge(ndim_, wpass);
        nwstop = lowpass_band_edge(ndim_, wstop);
        idx2 = nwpass - 1;
        idx3 = nwstop - 1;
    }
    Vec row(int k, double sign) const {
        Vec g(ndim);
        for (std::size_t j = 0; j < ndim; ++j) g[j] = sign > 0 ? spectrum[k * ndim + j] : -spectrum[k * ndim + j];
        return g;
    }
    FeasAnswer assess_feas(const Vec& x) {  // :58-131
        more_alt = true;
        for (int c = 0; c < nwpass; ++c) {
            idx1 += 1;
            if (idx1 == nwpass) idx1 = 0;
            const double val = dot(spectrum + idx1 * ndim, x.data(), ndim);
            if (val > up_sq) return CutPair{row(idx1, 1.0), Cut::parallel(val - up_sq, val - lp_sq)};
            if (val < lp_sq) return CutPair{row(idx1, -1.0), Cut::parallel(-val + lp_sq, -val + up_sq)};
        }
        fmax = -std::numeric_limits<double>::infinity();
        kmax = -1;
        for (int c = nwstop; c < static_cast<int>(mdim); ++c) {
            idx3 += 1;
            if (idx3 == static_cast<int>(mdim)) idx3 = nwstop;
            const double val = dot(spectrum + idx3 * ndim, x.data(), ndim);
            if (val > sp_sq) return CutPair{row(idx3, 1.0), Cut::parallel(val - sp_sq, val)};
            if (val < 0.0) return CutPair{row(idx3, -1.0), Cut::parallel(-val, -val + sp_sq)};
            if (val > fmax) { fmax = val; kmax = idx3; }
        }
        for (int c = nwpass; c < nwstop; ++c) {
            idx2 += 1;
            if (idx2 == nwstop) idx2 = nwpass;
            const double val = dot(spectrum + idx2 * ndim, x.data(), ndim);
            if (val < 0.0) return CutPair{row(idx2, -1.0), Cut::single(-val)};  // ParallelCut(-val, None)
        }
        more_alt = false;
        if (x[0] < 0.0) {
            Vec g(ndim, 0.0);
            g[0] = -1.0;
            return CutPair{g, Cut::single(-x[0])};
        }
        return std::nullopt;
    }
    OptimAnswer assess_optim(const Vec& x, double& sp) {  // :137-150
        sp_sq = sp;
        if (auto c = assess_feas(x)) return {*c, false};
        CutPair cut{row(kmax, 1.0), Cut::parallel(0.0, fmax)};
        sp = fmax;
        return {cut, true};
    }
};

// create_lowpass_case, :153-168 (degenerate: lp_sq > up_sq) — returns the 5 params
inline void lowpass_case_params(double out[5]) {
    const double PI = 3.14159265358979323846264338327950288;
    const double delta0_wpass = 0.025, delta0_wstop = 0.125;
    const double delta1 = 20.0 * std::log10(delta0_wpass * PI);
    const double delta2 = 20.0 * std::log10(delta0_wstop * PI);
    const double low_pass = std::pow(10.0, -delta1 / 20.0);
    const double up_pass = std::pow(10.0, delta1 / 20.0);
    const double stop_pass = std::pow(10.0, delta2 / 20.0);
    out[0] = 0.12; out[1] = 0.20;
    out[2] = low_pass * low_pass; out[3] = up_pass * up_pass; out[4] = stop_pass * stop_pass;
}

// ---- tests/integration_test.rs:90-105 (HighDimOracle: shifted quadratic bowl) ---------------------
struct Bowl {
    const double* target;
    std::size_t n;
    OptimAnswer assess_optim(const Vec& xc, double& gamma) {
        Vec diff(n), grad(n);
        for (std::size_t i = 0; i < n; ++i) diff[i] = xc[i] - target[i];
        for (std::size_t i = 0; i < n; ++i) grad[i] = 2.0 * diff[i];
        const double f = dot(diff, diff);
        if (f < gamma) {
            gamma = f;
            return {{grad, Cut::single(f)}, true};
        }
        return {{grad, Cut::single(f)}, false};
    }
};

// ---- src/oracles/ldlt_mgr.rs:3-124 ----------------------------------------------------------------
struct LDLTMgr {
    std::size_t pos0 = 0, pos1 = 0;
    Vec wit;
    std::size_t ndim;
    Vec storage;
    explicit LDLTMgr(std::size_t n) : wit(n, 0.0), ndim(n), storage(n * n, 0.0) {}
    template <class F>
    bool factor(F get_elem) {  // :29-55
        const std::size_t start = 0;
        pos0 = 0; pos1 = 0;
        for (std::size_t i = 0; i < ndim; ++i) {
            double diag = get_elem(i, start);
            for (std::size_t j = start; j < i; ++j) {
                storage[j * ndim + i] = diag;
                const double val = diag / storage[j * ndim + j];
                storage[i * ndim + j] = val;
                const std::size_t stop = j + 1;
                double s = 0.0;
                for (std::size_t k = start; k < stop; ++k) s += storage[i * ndim + k] * storage[k * ndim + stop];
                diag = get_elem(i, stop) - s;
            }
            storage[i * ndim + i] = diag;
            if (diag <= 0.0) {
                pos0 = start; pos1 = i + 1;
                break;
            }
        }
        return is_spd();
    }
    template <class F>
    bool factor_with_allow_semidefinite(F get_elem) {  // :59-89
        std::size_t start = 0;
        pos0 = 0; pos1 = 0;
        for (std::size_t i = 0; i < ndim; ++i) {
            double diag = get_elem(i, start);
            for (std::size_t j = start; j < i; ++j) {
                storage[j * ndim + i] = diag;
                const double val = diag / storage[j * ndim + j];
                storage[i * ndim + j] = val;
                const std::size_t stop = j + 1;
                double s = 0.0;
                for (std::size_t k = start; k < stop; ++k) s += storage[i * ndim + k] * storage[k * ndim + stop];
                diag = get_elem(i, stop) - s;
            }
            storage[i * ndim + i] = diag;
            if (diag < 0.0) {
                pos0 = start; pos1 = i + 1;
                break;
            } else if (diag == 0.0) {
                start = i + 1;
            }
        }
        return is_spd();
    }
    bool is_spd() const { return pos1 == 0; }
    double witness() {  // :98-111
        const std::size_t start = pos0, pos = pos1;
        const std::size_t m = pos - 1;
        wit[m] = 1.0;
        for (std::size_t i = m; i >= start + 1; --i) {
            double s = 0.0;
            for (std::size_t k = i; k < pos; ++k) s += storage[k * ndim + (i - 1)] * wit[k];
            wit[i - 1] = -s;
        }
        return -storage[m * ndim + m];
    }
    double sym_quad(const double* mat, std::size_t ld) const {  // :115-124
        double result = 0.0;
        for (std::size_t i = pos0; i < pos1; ++i)
            for (std::size_t j = pos0; j < pos1; ++j) result += wit[i] * mat[i * ld + j] * wit[j];
        return result;
    }
};

// ---- src/oracles/lmi_oracle.rs:5-45 ---------------------------------------------------------------
struct LMI {
    const double* mat_f;  // n matrices m×m, contiguous
    const double* mat_b;  // m×m
    std::size_t nvar, m;
    LDLTMgr mgr;
    LMI(const double* f, const double* b, std::size_t nvar_, std::size_t m_) : mat_f(f), mat_b(b), nvar(nvar_), m(m_), mgr(m_) {}
    FeasAnswer assess_feas(const Vec& xc) {
        const std::size_t n = xc.size();
        const bool feas = mgr.factor([&](std::size_t i, std::size_t j) {
            double s = mat_b[i * m + j];
            for (std::size_t k = 0; k < n; ++k) s -= mat_f[k * m * m + i * m + j] * xc[k];
            return s;
        });
        if (feas) return std::nullopt;
        const double ep = mgr.witness();
        Vec g(n);
        for (std::size_t k = 0; k < n; ++k) g[k] = mgr.sym_quad(mat_f + k * m * m, m);
        return CutPair{g, Cut::single(ep)};
    }
};

// ---- src/oracles/lmi_old_oracle.rs:5-50: forms the whole matrix A = B - sum_k F_k x_k (k outermost), then factorize ----
struct LMIOld {
    const double* mat_f;
    const double* mat_b;
    std::size_t nvar, m;
    LDLTMgr mgr;
    LMIOld(const double* f, const double* b, std::size_t nvar_, std::size_t m_) : mat_f(f), mat_b(b), nvar(nvar_), m(m_), mgr(m_) {}
    FeasAnswer assess_feas(const Vec& xc) {
        const std::size_t n = xc.size();
        Vec a(mat_b, mat_b + m * m);                                   // :30
        for (std::size_t k = 0; k < n; ++k) {                          // :31-40
            const double xk = xc[k];
            for (std::size_t i = 0; i < m; ++i)
                for (std::size_t j = 0; j < m; ++j) a[i * m + j] = a[i * m + j] - mat_f[k * m * m + i * m + j] * xk;
        }
        const bool feas = mgr.factor([&](std::size_t i, std::size_t j) { return a[i * m + j]; });   // factorize, ldlt_mgr.rs:22-24
        if (feas) return std::nullopt;
        const double ep = mgr.witness();
        Vec g(n);
        for (std::size_t k = 0; k < n; ++k) g[k] = mgr.sym_quad(mat_f + k * m * m, m);
        return CutPair{g, Cut::single(ep)};
    }
};

// ---- src/oracles/lmi0_oracle.rs:4-35: F(x) = sum_k F_k x_k > 0 (no constant term); the reference returns (g, ep) without
// a cut type — used as a feasibility oracle the pair becomes (g, SingleCut(ep)) like its two siblings ----
struct LMI0 {
    const double* mat_f;
    std::size_t nvar, m;
    LDLTMgr mgr;
    LMI0(const double* f, const double* /*unused*/, std::size_t nvar_, std::size_t m_) : mat_f(f), nvar(nvar_), m(m_), mgr(m_) {}
    FeasAnswer assess_feas(const Vec& x) {
        const std::size_t n = x.size();
        const bool feas = mgr.factor([&](std::size_t i, std::size_t j) {   // :18-24
            double s = 0.0;
            for (std::size_t k = 0; k < n; ++k) s += mat_f[k * m * m + i * m + j] * x[k];
            return s;
        });
        if (feas) return std::nullopt;
        const double ep = mgr.witness();
        Vec g(n);
        for (std::size_t k = 0; k < n; ++k) g[k] = -mgr.sym_quad(mat_f + k * m * m, m);   // :31
        return CutPair{g, Cut::single(ep)};
    }
};

// ---- tests/lmi_tests.rs:121-172 (MyLmiOracle<T> generalised to L blocks; L=2 in the reference) -------
template <class Block>
struct LmiLinOptT {
    int idx = -1;
    Vec c;
    std::vector<Block> blocks;
    OptimAnswer assess_optim(const Vec& xc, double& gamma) {
        const double f0 = dot(c, xc);
        const int nc = static_cast<int>(blocks.size()) + 1;
        for (int k = 0; k < nc; ++k) {
            idx = (idx == nc - 1) ? 0 : idx + 1;
            if (idx < nc - 1) {
                if (auto cut = blocks[idx].assess_feas(xc)) return {*cut, false};
            } else {
                const double fj = f0 - gamma;
                if (fj > 0.0) return {{c, Cut::single(fj)}, false};
                gamma = f0;
            }
        }
        return {{c, Cut::single(0.0)}, true};
    }
};
using LmiLinOpt = LmiLinOptT<LMI>;

// ---- tests/example2_tests.rs:12-48 ----------------------------------------------------------------
struct Example2 {
    int idx = -1;
    FeasAnswer assess_feas(const Vec& xc) {
        const double x = xc[0], y = xc[1];
        for (int k = 0; k < 2; ++k) {
            idx += 1;
            if (idx == 2) idx = 0;
            const double fj = idx == 0 ? x + y - 3.0 : -x + y + 1.0;
            if (fj > 0.0) return CutPair{idx == 0 ? Vec{1.0, 1.0} : Vec{-1.0, 1.0}, Cut::single(fj)};
        }
        return std::nullopt;
    }
};

// ---- tests/cutting_plane_tests.rs -----------------------------------------------------------------
struct HalfPlane {  // :12-28
    FeasAnswer assess_feas(const Vec& xc) {
        const double fj = xc[0] + xc[1] - 3.0;
        if (fj > 0.0) return CutPair{Vec{1.0, 1.0}, Cut::single(fj)};
        return std::nullopt;
    }
};
struct AlwaysCut {  // :33-42 (feas), :91-99 (optim), :104-122 (optim_q)
    FeasAnswer assess_feas(const Vec&) { return CutPair{Vec{1.0, 1.0}, Cut::single(1.0)}; }
    OptimAnswer assess_optim(const Vec&, double&) { return {{{1.0, 1.0}, Cut::single(1.0)}, false}; }
    QAnswer assess_optim_q(const Vec& xc, double&, bool) { return {{{1.0, 1.0}, Cut::single(1.0)}, false, xc, true}; }
};
struct OptimBox {  // :47-73
    OptimAnswer assess_optim(const Vec& xc, double& gamma) {
        const double x = xc[0], y = xc[1];
        const double f0 = x + y;
        const double f1 = x - 1.0;
        if (f1 > 0.0) return {{{1.0, 0.0}, Cut::single(f1)}, false};
        const double f2 = y - 1.0;
        if (f2 > 0.0) return {{{0.0, 1.0}, Cut::single(f2)}, false};
        const double f3 = f0 - gamma;
        if (f3 < 0.0) return {{{-1.0, -1.0}, Cut::single(-f3)}, false};
        return {{{-1.0, -1.0}, Cut::single(0.0)}, true};
    }
};
struct OptimBoxQ {  // :197-264 (MyOracleOptimQ3)
    QAnswer assess_optim_q(const Vec& xc, double& gamma, bool retry) {
        const double x = xc[0], y = xc[1];
        const double f0 = x + y;
        const double f1 = x - 1.0;
        if (f1 > 0.0) return {{{1.0, 0.0}, Cut::single(f1)}, false, xc, true};
        const double f2 = y - 1.0;
        if (f2 > 0.0) return {{{0.0, 1.0}, Cut::single(f2)}, false, xc, true};
        const double f3 = f0 - gamma;
        if (f3 < 0.0) return {{{-1.0, -1.0}, Cut::single(-f3)}, false, xc, true};
        const Vec xq{std::round(x), std::round(y)};
        const double f1q = xq[0] - 1.0;
        if (f1q > 0.0) return {{{1.0, 0.0}, Cut::single(f1q)}, false, xq, !retry};
        const double f2q = xq[1] - 1.0;
        if (f2q > 0.0) return {{{0.0, 1.0}, Cut::single(f2q)}, false, xq, !retry};
        const double f3q = xq[0] + xq[1] - gamma;
        if (f3q < 0.0) return {{{-1.0, -1.0}, Cut::single(-f3q)}, false, xq, !retry};
        gamma = xq[0] + xq[1];
        return {{{-1.0, -1.0}, Cut::single(0.0)}, true, xq, !retry};
    }
};
struct BS2 {  // :78-86
    bool assess_bs(double gamma) { return gamma > 0.0; }
};

}  // namespace ellref
