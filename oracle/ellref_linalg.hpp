// ellref_linalg.hpp — CPU restatement of the reference's GEMV-sharing numerics:
// src/power_iteration.rs:1-173 (four variants) and src/conjugate_gradient.rs:1-71.
// TEST INFRASTRUCTURE, NOT PRODUCT (see ellref.hpp).  `a` is n×n row-major; sums are sequential (src/arr.rs).
#pragma once
#include "ellref.hpp"

namespace ellref {

inline double norm_l1(const Vec& x) {  // power_iteration.rs:22-25
    double s = 0.0;
    for (double v : x) s += std::fabs(v);
    return s;
}
inline void div_assign(Vec& x, double s) {  // Arr /= f64 multiplies by the reciprocal (src/arr.rs:241-249)
    const double inv = 1.0 / s;
    for (double& v : x) v *= inv;
}
inline Vec mv(const Vec& a, const Vec& x) {
    Vec y(x.size());
    dot_mv(a.data(), x.data(), x.size(), x.size(), y.data());
    return y;
}
inline double l1_diff(const Vec& a, const Vec& b, double sign) {
    double s = 0.0;
    for (std::size_t i = 0; i < a.size(); ++i) s += std::fabs(a[i] + sign * b[i]);  // &*x - &x1 / &*x + &x1, then norm_l1
    return s;
}

inline std::pair<double, std::size_t> power_iteration(const Vec& a, Vec& x, const Options& o) {  // :53-65
    div_assign(x, std::sqrt(dot(x, x)));
    for (std::size_t niter = 0; niter < o.max_iters; ++niter) {
        const Vec x1 = x;
        x = mv(a, x1);
        div_assign(x, std::sqrt(dot(x, x)));
        if (l1_diff(x, x1, -1.0) <= o.tolerance || l1_diff(x, x1, 1.0) <= o.tolerance) return {dot(x, mv(a, x)), niter};
    }
    return {dot(x, mv(a, x)), o.max_iters};
}
inline std::pair<double, std::size_t> power_iteration4(const Vec& a, Vec& x, const Options& o) {  // :81-96
    div_assign(x, norm_l1(x));
    for (std::size_t niter = 0; niter < o.max_iters; ++niter) {
        const Vec x1 = x;
        x = mv(a, x1);
        div_assign(x, norm_l1(x));
        if (l1_diff(x, x1, -1.0) <= o.tolerance || l1_diff(x, x1, 1.0) <= o.tolerance) {
            div_assign(x, std::sqrt(dot(x, x)));
            return {dot(x, mv(a, x)), niter};
        }
    }
    div_assign(x, std::sqrt(dot(x, x)));
    return {dot(x, mv(a, x)), o.max_iters};
}
inline std::pair<double, std::size_t> power_iteration2(const Vec& a, Vec& x, const Options& o) {  // :112-127
    div_assign(x, std::sqrt(dot(x, x)));
    Vec new_vec = mv(a, x);
    double eigenval = dot(x, new_vec);
    for (std::size_t niter = 0; niter < o.max_iters; ++niter) {
        const double prev = eigenval;
        x = new_vec;
        div_assign(x, std::sqrt(dot(x, x)));
        new_vec = mv(a, x);
        eigenval = dot(x, new_vec);
        if (std::fabs(prev - eigenval) <= o.tolerance) return {eigenval, niter};
    }
    return {eigenval, o.max_iters};
}
inline std::pair<double, std::size_t> power_iteration3(const Vec& a, Vec& x, const Options& o) {  // :144-173
    Vec new_vec = mv(a, x);
    double d = dot(x, x);
    double eigenval = dot(x, new_vec) / d;
    for (std::size_t niter = 0; niter < o.max_iters; ++niter) {
        const double prev = eigenval;
        x = new_vec;
        d = dot(x, x);
        if (d >= 1e150) {
            div_assign(x, std::sqrt(dot(x, x)));
            new_vec = mv(a, x);
            eigenval = dot(x, new_vec);
            if (std::fabs(prev - eigenval) <= o.tolerance) return {eigenval, niter};
        } else {
            new_vec = mv(a, x);
            eigenval = dot(x, new_vec) / d;
            if (std::fabs(prev - eigenval) <= o.tolerance) {
                div_assign(x, std::sqrt(dot(x, x)));
                return {eigenval, niter};
            }
        }
    }
    return {eigenval, o.max_iters};
}

// conjugate_gradient.rs:21-71: returns iterations used (>= 0) on Ok, -1 "did not converge" (zero curvature),
// -2 "did not converge after max iterations"; x holds the iterate in every case
inline long conjugate_gradient(const Vec& a, const Vec& b, Vec& x, double tol, std::size_t max_iter) {
    const std::size_t n = b.size();
    Vec ax = mv(a, x);
    Vec residual(n);
    for (std::size_t i = 0; i < n; ++i) residual[i] = b[i] - ax[i];
    Vec direction = residual;
    double rns = dot(residual, residual);
    for (std::size_t it = 0; it < max_iter; ++it) {
        const Vec a_dir = mv(a, direction);
        const double dad = dot(direction, a_dir);
        if (dad == 0.0) return -1;
        const double step = rns / dad;
        for (std::size_t i = 0; i < n; ++i) x[i] += step * direction[i];
        for (std::size_t i = 0; i < n; ++i) residual[i] -= step * a_dir[i];
        const double rns_new = dot(residual, residual);
        if (std::sqrt(rns_new) < tol) return static_cast<long>(it);
        const double ratio = rns_new / rns;
        for (std::size_t i = 0; i < n; ++i) direction[i] = residual[i] + ratio * direction[i];
        rns = rns_new;
    }
    return -2;
}

}  // namespace ellref
