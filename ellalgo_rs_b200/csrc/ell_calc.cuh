// ell_calc.cuh — device-side cut calculus: (beta, tsq) -> (status, rho, sigma, delta).
//
// Follows the reference's EllCalc / EllCalcCore (src/ell_calc.rs:61-78 constants, :218-240 parallel deep cut,
// :383-394 parallel central cut, :453-459 / :550-553 deep cut, :605-611 central cut, guards :671-931).
// The whole library is compiled with -fmad=false, so every a*b+c below is a separately rounded multiply and add,
// exactly like the Rust code; expressions keep the reference's association.
#pragma once
#include <cuda_runtime.h>

namespace ellcuda {

enum : int { kSuccess = 0, kNoSoln = 1, kNoEffect = 2, kUnknown = 3 };
enum : int { kExitNoSoln = 1, kExitNoEffect = 2, kExitTol = 4, kExitMaxIt = 5, kExitFeasible = 6 };
enum : int { kBias = 0, kCentral = 1, kQ = 2 };

// SingleCut(beta0) / ParallelCut(beta0, None)  -> has1 = false;  ParallelCut(beta0, Some(beta1)) -> has1 = true.
// (None dispatches to the single-cut formulas, src/ell_calc.rs:671-718.)
struct Cut {
    double b0;
    double b1;
    bool has1;
};
__host__ __device__ inline Cut single_cut(double b) { return Cut{b, 0.0, false}; }
__host__ __device__ inline Cut parallel_cut(double b0, double b1) { return Cut{b0, b1, true}; }

struct CalcConsts {  // EllCalcCore::new, src/ell_calc.rs:61-78
    double n_f, n_plus_1, half_n, inv_n, cst1, cst2;
    int use_parallel_cut;
    __host__ __device__ static CalcConsts make(int n, int use_parallel) {
        CalcConsts c;
        c.n_f = (double)n;
        c.n_plus_1 = c.n_f + 1.0;
        c.half_n = c.n_f / 2.0;
        c.inv_n = 1.0 / c.n_f;
        const double n_sq = c.n_f * c.n_f;
        const double cst0 = 1.0 / (c.n_f + 1.0);
        c.cst1 = n_sq / (n_sq - 1.0);
        c.cst2 = 2.0 * cst0;
        c.use_parallel_cut = use_parallel;
        return c;
    }
};

struct Rsd {
    double rho, sigma, delta;
};

__host__ __device__ inline Rsd deep_cut_fast(const CalcConsts& c, double beta, double tau, double eta) {  // :453-459
    Rsd r;
    r.rho = eta / c.n_plus_1;
    r.sigma = 2.0 * r.rho / (tau + beta);
    const double alpha = beta / tau;
    r.delta = c.cst1 * (1.0 - alpha * alpha);
    return r;
}

__host__ __device__ inline Rsd parallel_deep_cut_fast(const CalcConsts& c, double b0, double b1, double tsq, double b0b1,
                                                      double eta) {  // :218-240
    const double b0sq = b0 * b0;
    const double b1sq = b1 * b1;
    const double zeta0 = tsq - b0sq;
    const double zeta1 = tsq - b1sq;
    const double temp = c.half_n * (b1sq - b0sq);
    const double xi = sqrt(zeta0 * zeta1 + temp * temp);
    const double bsum = b0 + b1;
    const double bsumsq = bsum * bsum;
    Rsd r;
    r.sigma = 2.0 * eta / (tsq + b0b1 + c.half_n * bsumsq + xi);
    r.rho = r.sigma * (b0 + b1) / 2.0;
    r.delta = c.cst1 * ((zeta0 + zeta1) / 2.0 + xi / c.n_f) / tsq;
    return r;
}

__host__ __device__ inline Rsd parallel_central_cut_core(const CalcConsts& c, double b1, double tsq) {  // :383-394
    const double b1sq = b1 * b1;
    const double a1sq = b1sq / tsq;
    const double half_val = c.half_n * a1sq;
    const double root = half_val + sqrt(1.0 - a1sq + half_val * half_val);
    const double r_plus_1 = root + 1.0;
    Rsd r;
    r.rho = b1 / r_plus_1;
    r.sigma = 2.0 / r_plus_1;
    r.delta = root / (root - c.inv_n);
    return r;
}

__host__ __device__ inline int central_cut(const CalcConsts& c, double tsq, Rsd& r) {  // :605-611, :928-931
    r.sigma = c.cst2;
    r.rho = sqrt(tsq) / c.n_plus_1;
    r.delta = c.cst1;
    return kSuccess;
}

__host__ __device__ inline int deep_cut(const CalcConsts& c, double beta, double tsq, Rsd& r) {  // :870-877
    if (tsq < beta * beta) {
        r = Rsd{0.0, 0.0, 0.0};
        return kNoSoln;
    }
    const double tau = sqrt(tsq);
    const double eta = tau + c.n_f * beta;  // :550-553
    r = deep_cut_fast(c, beta, tau, eta);
    return kSuccess;
}

__host__ __device__ inline int deep_cut_q(const CalcConsts& c, double beta, double tsq, Rsd& r) {  // :892-908
    const double tau = sqrt(tsq);
    if (tau < beta) {
        r = Rsd{0.0, 0.0, 0.0};
        return kNoSoln;
    }
    const double eta = tau + c.n_f * beta;
    if (eta < 0.0) {
        r = Rsd{0.0, 0.0, 1.0};
        return kNoEffect;
    }
    r = deep_cut_fast(c, beta, tau, eta);
    return kSuccess;
}

// The three CutType closures of src/ell.rs:182-229 / src/ell_stable.rs:168-210 in one dispatcher.
__host__ __device__ inline int ell_calc(const CalcConsts& c, int which, const Cut& cut, double tsq, Rsd& r) {
    if (!cut.has1) {
        if (which == kBias) return deep_cut(c, cut.b0, tsq, r);
        if (which == kCentral) return central_cut(c, tsq, r);
        return deep_cut_q(c, cut.b0, tsq, r);
    }
    const double b0 = cut.b0, b1 = cut.b1;
    if (which == kBias) {  // calc_parallel_bias_cut :751-769
        if (b1 < b0) {
            r = Rsd{0.0, 0.0, 0.0};
            return kNoSoln;
        }
        if ((b1 > 0.0 && tsq <= b1 * b1) || !c.use_parallel_cut) return deep_cut(c, b0, tsq, r);
        const double b0b1 = b0 * b1;  // :316-320
        const double eta = tsq + c.n_f * b0b1;
        r = parallel_deep_cut_fast(c, b0, b1, tsq, b0b1, eta);
        return kSuccess;
    }
    if (which == kCentral) {  // calc_parallel_central_cut :836-847 (beta0 is ignored, :694)
        if (b1 < 0.0) {
            r = Rsd{0.0, 0.0, 0.0};
            return kNoSoln;
        }
        if (tsq <= b1 * b1 || !c.use_parallel_cut) return central_cut(c, tsq, r);
        r = parallel_central_cut_core(c, b1, tsq);
        return kSuccess;
    }
    // calc_parallel_q :787-812
    if (b1 < b0) {
        r = Rsd{0.0, 0.0, 0.0};
        return kNoSoln;
    }
    if (((b1 > 0.0) && b1 * b1 >= tsq) || !c.use_parallel_cut) return deep_cut_q(c, b0, tsq, r);
    const double b0b1 = b0 * b1;
    const double eta = tsq + c.n_f * b0b1;
    if (eta <= 0.0) {
        r = Rsd{0.0, 0.0, 1.0};
        return kNoEffect;
    }
    r = parallel_deep_cut_fast(c, b0, b1, tsq, b0b1, eta);
    return kSuccess;
}

}  // namespace ellcuda
