// ellref.hpp — CPU restatement of the reference ellipsoid path (luk036/ellalgo-rs v0.1.7).
//
// TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference leg may build or call this.  The product (libellcuda.so)
// never links or calls it.
//
// The Rust reference cannot be compiled in this image (no cargo/rustc), so this is a
// line-by-line restatement; every function names the reference lines it follows
// (paths relative to the reference root).  Numeric rules: IEEE f64,
// no FMA contraction (build with -ffp-contract=off), no reassociation, sums sequential from
// 0.0, libm = glibc (what Rust's f64::exp/ln/cos/sqrt call on Linux).
//
// Pinning: tests/test_oracle_pins.py checks this file against every iteration count and
// known-answer vector the reference's own tests hold for the path.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <optional>
#include <utility>
#include <vector>

namespace ellref {

using Vec = std::vector<double>;

// ---- src/cutting_plane.rs:32-37 -----------------------------------------------------------
enum CutStatus : int { Success = 0, NoSoln = 1, NoEffect = 2, Unknown = 3 };

// why a driver returned (not part of the reference's return value; see include/ellcuda.h)
enum ExitStatus : int { ExitNoSoln = 1, ExitNoEffect = 2, ExitTol = 4, ExitMaxIt = 5, ExitFeasible = 6 };

// SingleCut(beta) (src/cutting_plane.rs:9) and ParallelCut(beta0, Option<beta1>) (:18) in one
// record.  A ParallelCut with None dispatches to exactly the single-cut formulas
// (src/ell_calc.rs:671-718), so `has_beta1 == false` covers both.
struct Cut {
    double beta0 = 0.0;
    bool has_beta1 = false;
    double beta1 = 0.0;
    static Cut single(double b) { return Cut{b, false, 0.0}; }
    static Cut parallel(double b0, double b1) { return Cut{b0, true, b1}; }
};

struct Options {  // src/cutting_plane.rs:50-100
    std::size_t max_iters = 2000;
    double tolerance = 1e-20;
};

// ---- src/arr.rs -----------------------------------------------------------------------------
inline double dot(const double* a, const double* b, std::size_t n) {  // arr.rs:443-451
    double s = 0.0;
    for (std::size_t k = 0; k < n; ++k) s += a[k] * b[k];
    return s;
}
inline double dot(const Vec& a, const Vec& b) { return dot(a.data(), b.data(), a.size()); }

inline void dot_mv(const double* m, const double* x, std::size_t rows, std::size_t cols, double* out) {
    for (std::size_t i = 0; i < rows; ++i) {  // arr.rs:426-442
        double acc = 0.0;
        const double* row = m + i * cols;
        for (std::size_t j = 0; j < cols; ++j) acc += row[j] * x[j];
        out[i] = acc;
    }
}

inline Vec linspace(double start, double end, std::size_t n) {  // arr.rs:506-520
    Vec a(n, 0.0);
    if (n == 0) return a;
    if (n == 1) { a[0] = start; return a; }
    const double step = (end - start) / static_cast<double>(n - 1);
    for (std::size_t i = 0; i < n; ++i) a[i] = start + step * static_cast<double>(i);
    return a;
}

// ---- src/ell_calc.rs ---------------------------------------------------------------------------
struct Rsd { double rho, sigma, delta; };
struct CalcResult { CutStatus status; Rsd v; };

struct EllCalcCore {  // ell_calc.rs:22-78
    double n_f, n_plus_1, half_n, inv_n, cst1, cst2;
    explicit EllCalcCore(double nf) {
        n_f = nf;
        n_plus_1 = nf + 1.0;
        half_n = nf / 2.0;
        inv_n = 1.0 / nf;
        const double n_sq = nf * nf;
        const double cst0 = 1.0 / (nf + 1.0);
        cst1 = n_sq / (n_sq - 1.0);
        cst2 = 2.0 * cst0;
    }
    Rsd calc_parallel_bias_cut_fast(double b0, double b1, double tsq, double b0b1, double eta) const {
        const double b0sq = b0 * b0;  // ell_calc.rs:218-240
        const double b1sq = b1 * b1;
        const double zeta0 = tsq - b0sq;
        const double zeta1 = tsq - b1sq;
        const double temp = half_n * (b1sq - b0sq);
        const double xi = std::sqrt(zeta0 * zeta1 + temp * temp);
        const double bsum = b0 + b1;
        const double bsumsq = bsum * bsum;
        const double sigma = 2.0 * eta / (tsq + b0b1 + half_n * bsumsq + xi);
        const double rho = sigma * (b0 + b1) / 2.0;
        const double delta = cst1 * ((zeta0 + zeta1) / 2.0 + xi / n_f) / tsq;
        return {rho, sigma, delta};
    }
    Rsd calc_parallel_bias_cut(double b0, double b1, double tsq) const {  // :316-320
        const double b0b1 = b0 * b1;
        const double eta = tsq + n_f * b0b1;
        return calc_parallel_bias_cut_fast(b0, b1, tsq, b0b1, eta);
    }
    Rsd calc_parallel_central_cut(double b1, double tsq) const {  // :383-394
        const double b1sq = b1 * b1;
        const double a1sq = b1sq / tsq;
        const double half_val = half_n * a1sq;
        const double root = half_val + std::sqrt(1.0 - a1sq + half_val * half_val);
        const double r_plus_1 = root + 1.0;
        const double rho = b1 / r_plus_1;
        const double sigma = 2.0 / r_plus_1;
        const double delta = root / (root - inv_n);
        return {rho, sigma, delta};
    }
    Rsd calc_bias_cut_fast(double beta, double tau, double eta) const {  // :453-459
        const double rho = eta / n_plus_1;
        const double sigma = 2.0 * rho / (tau + beta);
        const double alpha = beta / tau;
        const double delta = cst1 * (1.0 - alpha * alpha);
        return {rho, sigma, delta};
    }
    Rsd calc_bias_cut(double beta, double tau) const {  // :550-553
        const double eta = tau + n_f * beta;
        return calc_bias_cut_fast(beta, tau, eta);
    }
    Rsd calc_central_cut(double tsq) const {  // :605-611
        const double sigma = cst2;
        const double rho = std::sqrt(tsq) / n_plus_1;
        const double delta = cst1;
        return {rho, sigma, delta};
    }
};

struct EllCalc {  // ell_calc.rs:627-931
    double n_f;
    EllCalcCore helper;
    bool use_parallel_cut = true;
    explicit EllCalc(std::size_t n) : n_f(static_cast<double>(n)), helper(static_cast<double>(n)) {}

    CalcResult calc_central_cut(double tsq) const { return {Success, helper.calc_central_cut(tsq)}; }  // :928-931
    CalcResult calc_bias_cut(double beta, double tsq) const {  // :870-877
        if (tsq < beta * beta) return {NoSoln, {0.0, 0.0, 0.0}};
        const double tau = std::sqrt(tsq);
        return {Success, helper.calc_bias_cut(beta, tau)};
    }
    CalcResult calc_bias_cut_q(double beta, double tsq) const {  // :892-908
        const double tau = std::sqrt(tsq);
        if (tau < beta) return {NoSoln, {0.0, 0.0, 0.0}};
        const double eta = tau + n_f * beta;
        if (eta < 0.0) return {NoEffect, {0.0, 0.0, 1.0}};
        return {Success, helper.calc_bias_cut_fast(beta, tau, eta)};
    }
    CalcResult calc_parallel_bias_cut(double b0, double b1, double tsq) const {  // :751-769
        if (b1 < b0) return {NoSoln, {0.0, 0.0, 0.0}};
        if ((b1 > 0.0 && tsq <= b1 * b1) || !use_parallel_cut) return calc_bias_cut(b0, tsq);
        return {Success, helper.calc_parallel_bias_cut(b0, b1, tsq)};
    }
    CalcResult calc_parallel_q(double b0, double b1, double tsq) const {  // :787-812
        if (b1 < b0) return {NoSoln, {0.0, 0.0, 0.0}};
        if (((b1 > 0.0) && b1 * b1 >= tsq) || !use_parallel_cut) return calc_bias_cut_q(b0, tsq);
        const double b0b1 = b0 * b1;
        const double eta = tsq + n_f * b0b1;
        if (eta <= 0.0) return {NoEffect, {0.0, 0.0, 1.0}};
        return {Success, helper.calc_parallel_bias_cut_fast(b0, b1, tsq, b0b1, eta)};
    }
    CalcResult calc_parallel_central_cut(double b1, double tsq) const {  // :836-847
        if (b1 < 0.0) return {NoSoln, {0.0, 0.0, 0.0}};
        if (tsq <= b1 * b1 || !use_parallel_cut) return calc_central_cut(tsq);
        return {Success, helper.calc_parallel_central_cut(b1, tsq)};
    }
    // :671-718 (the Option<beta1> dispatch)
    CalcResult bias(const Cut& c, double tsq) const {
        return c.has_beta1 ? calc_parallel_bias_cut(c.beta0, c.beta1, tsq) : calc_bias_cut(c.beta0, tsq);
    }
    CalcResult central(const Cut& c, double tsq) const {
        return c.has_beta1 ? calc_parallel_central_cut(c.beta1, tsq) : calc_central_cut(tsq);
    }
    CalcResult q(const Cut& c, double tsq) const {
        return c.has_beta1 ? calc_parallel_q(c.beta0, c.beta1, tsq) : calc_bias_cut_q(c.beta0, tsq);
    }
    CalcResult dispatch(int which, const Cut& c, double tsq) const {
        return which == 0 ? bias(c, tsq) : (which == 1 ? central(c, tsq) : q(c, tsq));
    }
};

// ---- src/ell.rs ------------------------------------------------------------------------------------
struct Ell {
    bool no_defer_trick = false;
    std::size_t n;
    Vec mq;  // n×n row-major, both triangles
    Vec xc;
    double kappa;
    EllCalc helper;
    double tsq = 0.0;

    Ell(double kappa_, Vec mq_, Vec xc_)  // new_with_matrix, ell.rs:31-41
        : n(xc_.size()), mq(std::move(mq_)), xc(std::move(xc_)), kappa(kappa_), helper(n) {}
    static Ell with_diag(const Vec& val, Vec xc_) {  // Ell::new, ell.rs:54-56
        const std::size_t n = xc_.size();
        Vec m(n * n, 0.0);
        for (std::size_t i = 0; i < n; ++i) m[i * n + i] = val[i];
        return Ell(1.0, std::move(m), std::move(xc_));
    }
    static Ell with_scalar(double val, Vec xc_) {  // new_with_scalar, ell.rs:65-67
        const std::size_t n = xc_.size();
        Vec m(n * n, 0.0);
        for (std::size_t i = 0; i < n; ++i) m[i * n + i] = 1.0;
        return Ell(val, std::move(m), std::move(xc_));
    }

    // update_core, ell.rs:97-137.  which: 0 bias, 1 central, 2 q (the three CutType closures :215-228)
    CutStatus update(int which, const Vec& grad, const Cut& cut) {
        Vec grad_t(n);
        dot_mv(mq.data(), grad.data(), n, n, grad_t.data());
        const double omega = dot(grad, grad_t);
        tsq = kappa * omega;
        const CalcResult r = helper.dispatch(which, cut, tsq);
        if (r.status != Success) return r.status;
        const double rho_over_omega = r.v.rho / omega;
        for (std::size_t i = 0; i < n; ++i) xc[i] -= rho_over_omega * grad_t[i];
        const double ratio = r.v.sigma / omega;
        for (std::size_t i = 0; i < n; ++i) {
            const double r_qg = ratio * grad_t[i];
            for (std::size_t j = 0; j <= i; ++j) {
                const double update = r_qg * grad_t[j];
                const std::size_t idx = i * n + j;
                mq[idx] -= update;
                if (i != j) mq[j * n + i] = mq[idx];
            }
        }
        kappa *= r.v.delta;
        if (no_defer_trick) {
            for (double& v : mq) v *= kappa;
            kappa = 1.0;
        }
        return r.status;
    }
    void set_use_parallel_cut(bool b) { helper.use_parallel_cut = b; }
};

// ---- src/ell_stable.rs ----------------------------------------------------------------------------
struct EllStable {
    std::size_t n;
    Vec mq;
    Vec xc;
    double kappa;
    EllCalc helper;
    double tsq = 0.0;

    EllStable(double kappa_, Vec mq_, Vec xc_)  // ell_stable.rs:18-27
        : n(xc_.size()), mq(std::move(mq_)), xc(std::move(xc_)), kappa(kappa_), helper(n) {}
    static EllStable with_diag(const Vec& val, Vec xc_) {  // :29-31
        const std::size_t n = xc_.size();
        Vec m(n * n, 0.0);
        for (std::size_t i = 0; i < n; ++i) m[i * n + i] = val[i];
        return EllStable(1.0, std::move(m), std::move(xc_));
    }
    static EllStable with_scalar(double val, Vec xc_) {  // :33-35
        const std::size_t n = xc_.size();
        Vec m(n * n, 0.0);
        for (std::size_t i = 0; i < n; ++i) m[i * n + i] = 1.0;
        return EllStable(val, std::move(m), std::move(xc_));
    }

    // update_core, ell_stable.rs:52-125
    CutStatus update(int which, const Vec& grad, const Cut& cut) {
        const std::size_t ndim = n;
        Vec inv_ml_g = grad;  // :60-69
        for (std::size_t i = 1; i < ndim; ++i) {
            for (std::size_t j = 0; j < i; ++j) {
                const double val = mq[j * n + i] * inv_ml_g[j];
                mq[i * n + j] = val;
                inv_ml_g[i] -= val;
            }
        }
        Vec inv_md_inv_ml_g = inv_ml_g;  // :72-75
        for (std::size_t i = 0; i < ndim; ++i) inv_md_inv_ml_g[i] *= mq[i * n + i];
        Vec gg_t = inv_md_inv_ml_g;  // :78-83
        double omega = 0.0;
        for (std::size_t i = 0; i < ndim; ++i) {
            gg_t[i] *= inv_ml_g[i];
            omega += gg_t[i];
        }
        tsq = kappa * omega;  // :85-90
        const CalcResult r = helper.dispatch(which, cut, tsq);
        if (r.status != Success) return r.status;
        Vec g_t = inv_md_inv_ml_g;  // :93-98
        for (std::size_t i = ndim - 1; i >= 1; --i) {
            for (std::size_t j = i; j < ndim; ++j) g_t[i - 1] -= mq[j * n + (i - 1)] * g_t[j];
        }
        const double rho_over_omega = r.v.rho / omega;  // :101-104
        for (std::size_t i = 0; i < ndim; ++i) xc[i] -= rho_over_omega * g_t[i];
        const double mu_val = r.v.sigma / (1.0 - r.v.sigma);  // :107-121
        double oldt = omega / mu_val;
        const std::size_t last_idx = ndim - 1;
        for (std::size_t j = 0; j < last_idx; ++j) {
            const double temp = oldt + gg_t[j];
            const double beta2 = inv_md_inv_ml_g[j] / temp;
            mq[j * n + j] *= oldt / temp;
            for (std::size_t l = j + 1; l < ndim; ++l) mq[j * n + l] += beta2 * mq[l * n + j];
            oldt = temp;
        }
        const double temp = oldt + gg_t[last_idx];
        mq[last_idx * n + last_idx] *= oldt / temp;
        kappa *= r.v.delta;  // :122
        return r.status;
    }
    void set_use_parallel_cut(bool b) { helper.use_parallel_cut = b; }
};

// ---- drivers, src/cutting_plane.rs ---------------------------------------------------------------
struct RunResult {
    std::optional<Vec> xbest;
    std::size_t num_iters = 0;
    int exit_status = 0;
};

// Oracle concept for feas: std::optional<std::pair<Vec,Cut>> assess_feas(const Vec& xc)
template <class Oracle, class Space>
RunResult cutting_plane_feas(Oracle& omega, Space& space, const Options& opt) {  // :205-227
    for (std::size_t niter = 0; niter < opt.max_iters; ++niter) {
        auto cut = omega.assess_feas(space.xc);
        if (!cut) return {space.xc, niter, ExitFeasible};
        const CutStatus status = space.update(0, cut->first, cut->second);
        if (status != Success) return {std::nullopt, niter, static_cast<int>(status)};
        if (space.tsq < opt.tolerance) return {std::nullopt, niter, ExitTol};
    }
    return {std::nullopt, opt.max_iters, ExitMaxIt};
}

// Oracle concept for optim: std::pair<std::pair<Vec,Cut>,bool> assess_optim(const Vec& xc, double& gamma)
template <class Oracle, class Space>
RunResult cutting_plane_optim(Oracle& omega, Space& space, double& gamma, const Options& opt) {  // :286-313
    std::optional<Vec> x_best;
    for (std::size_t niter = 0; niter < opt.max_iters; ++niter) {
        auto [cut, shrunk] = omega.assess_optim(space.xc, gamma);
        CutStatus status;
        if (shrunk) {
            x_best = space.xc;
            status = space.update(1, cut.first, cut.second);
        } else {
            status = space.update(0, cut.first, cut.second);
        }
        if (status != Success) return {x_best, niter, static_cast<int>(status)};
        if (space.tsq < opt.tolerance) return {x_best, niter, ExitTol};
    }
    return {x_best, opt.max_iters, ExitMaxIt};
}

struct QAnswer {
    std::pair<Vec, Cut> cut;
    bool shrunk;
    Vec x_q;
    bool more_alt;
};
// Oracle concept for optim_q: QAnswer assess_optim_q(const Vec& xc, double& gamma, bool retry)
template <class Oracle, class Space>
RunResult cutting_plane_optim_q(Oracle& omega, Space& space, double& gamma, const Options& opt) {  // :331-374
    std::optional<Vec> x_best;
    bool retry = false;
    for (std::size_t niter = 0; niter < opt.max_iters; ++niter) {
        QAnswer a = omega.assess_optim_q(space.xc, gamma, retry);
        if (a.shrunk) {
            x_best = a.x_q;
            retry = false;
        }
        const CutStatus status = space.update(2, a.cut.first, a.cut.second);
        if (status == Success) {
            retry = false;
        } else if (status == NoSoln) {
            return {x_best, niter, ExitNoSoln};
        } else if (status == NoEffect) {
            if (!a.more_alt) return {x_best, niter, ExitNoEffect};
            retry = true;
        }
        if (space.tsq < opt.tolerance) return {x_best, niter, ExitTol};
    }
    return {x_best, opt.max_iters, ExitMaxIt};
}

// bsearch + BSearchAdaptor, src/cutting_plane.rs:376-466
template <class Oracle, class Space>
struct BSearchAdaptor {
    Oracle omega;
    Space space;
    Options options;
    bool assess_bs(double gamma) {  // :409-418
        Space sp = space;
        omega.update(gamma);
        RunResult r = cutting_plane_feas(omega, sp, options);
        if (r.xbest) {
            space.xc = *r.xbest;
            return true;
        }
        return false;
    }
};
// `final_upper` (not in the reference's signature) reports the interval's upper end on return, for parity checks.
template <class OracleBS>
std::pair<bool, std::size_t> bsearch(OracleBS& omega, std::pair<double, double> intrvl, const Options& opt,
                                     double* final_upper = nullptr) {  // :441-466
    double lower = intrvl.first, upper = intrvl.second;
    const double u_orig = upper;
    for (std::size_t niter = 0; niter < opt.max_iters; ++niter) {
        const double tau = (upper - lower) / 2.0;
        if (tau < opt.tolerance) {
            if (final_upper) *final_upper = upper;
            return {upper != u_orig, niter};
        }
        double gamma = lower;
        gamma += tau;
        if (omega.assess_bs(gamma)) upper = gamma; else lower = gamma;
    }
    if (final_upper) *final_upper = upper;
    return {upper != u_orig, opt.max_iters};
}

}  // namespace ellref
