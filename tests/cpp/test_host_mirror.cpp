// test_host_mirror.cpp — the reference's own driver tests, run through include/ellalgo_cuda.hpp (C++ host mirror over
// the C ABI) with the search space on the GPU.  Mirrors src/example1.rs:39-70, src/example3.rs:68-85,
// src/oracles/profit_oracle.rs:207-223 and 244-260, tests/cutting_plane_tests.rs:128-146.
// Usage: test_host_mirror            -> runs on cuda:0, exit 0 on success
//        test_host_mirror --no-gpu   -> checks that construction fails loudly without a device (no CPU fallback)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>

#include "ellalgo_cuda.hpp"

using namespace ellalgo;

#define EXPECT(cond)                                                                  \
    do {                                                                              \
        if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
    } while (0)

struct Example1 {  // src/example1.rs:8-30
    std::pair<std::pair<Arr, SingleCut>, bool> assess_optim(const Arr& xc, double& gamma) {
        const double x = xc[0], y = xc[1];
        const double f0 = x + y;
        const double f1 = f0 - 3.0;
        if (f1 > 0.0) return {{{1.0, 1.0}, SingleCut{f1}}, false};
        const double f2 = -x + y + 1.0;
        if (f2 > 0.0) return {{{-1.0, 1.0}, SingleCut{f2}}, false};
        const double f3 = gamma - f0;
        if (f3 > 0.0) return {{{-1.0, -1.0}, SingleCut{f3}}, false};
        gamma = f0;
        return {{{-1.0, -1.0}, SingleCut{0.0}}, true};
    }
};

struct Example3 {  // src/example3.rs:10-60
    int idx = -1;
    double target = -1e100;
    std::optional<std::pair<Arr, SingleCut>> assess_feas(const Arr& xc) {
        const double x = xc[0], y = xc[1];
        for (int k = 0; k < 4; ++k) {
            idx = idx == 3 ? 0 : idx + 1;
            const double fj = idx == 0 ? -x - 1.0 : idx == 1 ? -y - 2.0 : idx == 2 ? x + y - 1.0 : 2.0 * x - 3.0 * y - target;
            if (fj > 0.0) {
                const Arr g = idx == 0 ? Arr{-1.0, 0.0} : idx == 1 ? Arr{0.0, -1.0} : idx == 2 ? Arr{1.0, 1.0} : Arr{2.0, -3.0};
                return std::make_pair(g, SingleCut{fj});
            }
        }
        return std::nullopt;
    }
    void update(double gamma) { target = gamma; }
};

struct Profit {  // src/oracles/profit_oracle.rs:42-114
    int idx = -1;
    double log_p_scale, log_k, v[2], a[2], log_cobb = 0, vx = 0, q[2] = {0, 0};
    Profit(double p, double scale, double limit, double a0, double a1, double v0, double v1)
        : log_p_scale(std::log(p * scale)), log_k(std::log(limit)), v{v0, v1}, a{a0, a1} {}
    std::optional<std::pair<Arr, double>> feas(const Arr& y, double gamma) {
        for (int k = 0; k < 2; ++k) {
            idx = idx == 1 ? 0 : idx + 1;
            double fj;
            if (idx == 0) fj = y[0] - log_k;
            else {
                double d = 0.0; d += a[0] * y[0]; d += a[1] * y[1];
                log_cobb = log_p_scale + d;
                q[0] = v[0] * std::exp(y[0]); q[1] = v[1] * std::exp(y[1]);
                vx = q[0] + q[1];
                fj = std::log(gamma + vx) - log_cobb;
            }
            if (fj > 0.0) {
                if (idx == 0) return std::make_pair(Arr{1.0, 0.0}, fj);
                const double inv = 1.0 / (gamma + vx);
                return std::make_pair(Arr{q[0] * inv - a[0], q[1] * inv - a[1]}, fj);
            }
        }
        return std::nullopt;
    }
    std::pair<std::pair<Arr, SingleCut>, bool> assess_optim(const Arr& y, double& gamma) {
        if (auto c = feas(y, gamma)) return {{c->first, SingleCut{c->second}}, false};
        const double e = std::exp(log_cobb);
        gamma = e - vx;
        const double inv = 1.0 / e;
        return {{{q[0] * inv - a[0], q[1] * inv - a[1]}, SingleCut{0.0}}, true};
    }
};

struct AlwaysCut {  // tests/cutting_plane_tests.rs:33-42
    std::optional<std::pair<Arr, SingleCut>> assess_feas(const Arr&) { return std::make_pair(Arr{1.0, 1.0}, SingleCut{1.0}); }
};

int main(int argc, char** argv) {
    if (argc > 1 && !std::strcmp(argv[1], "--no-gpu")) {
        try {
            auto e = EllCuda::new_with_scalar(10.0, {0.0, 0.0});
            std::printf("FAILED: construction succeeded without a GPU\n");
            return 1;
        } catch (const CudaError& err) {
            std::printf("ok: %s\n", err.what());
            return err.code == ELLCUDA_E_CUDA ? 0 : 1;
        }
    }
    const double NEG_INF = -std::numeric_limits<double>::infinity();
    {   // example1::test_feasible
        auto ellip = EllCuda::new_with_scalar(10.0, {0.0, 0.0});
        Example1 oracle;
        double gamma = NEG_INF;
        Options options; options.tolerance = 1e-10;
        auto [xbest, num_iters] = cutting_plane_optim(oracle, ellip, gamma, options);
        EXPECT(xbest.has_value());
        EXPECT(num_iters == 25);
    }
    {   // example1::test_infeasible1 / 2
        auto ellip = EllCuda::new_({10.0, 10.0}, {100.0, 100.0});
        Example1 oracle;
        double gamma = NEG_INF;
        EXPECT(!cutting_plane_optim(oracle, ellip, gamma, Options{}).first.has_value());
        auto ellip2 = EllCuda::new_({10.0, 10.0}, {0.0, 0.0});
        double g2 = 100.0;
        EXPECT(!cutting_plane_optim(oracle, ellip2, g2, Options{}).first.has_value());
    }
    {   // profit_oracle::test_profit_oracle
        auto ellip = EllCuda::new_({100.0, 100.0}, {0.0, 0.0});
        Profit omega(20.0, 40.0, 30.5, 0.1, 0.4, 10.0, 35.0);
        double gamma = 0.0;
        auto [y, niter] = cutting_plane_optim(omega, ellip, gamma, Options{});
        EXPECT(y.has_value());
        EXPECT((*y)[0] <= std::log(30.5));
        EXPECT(niter == 83);
    }
    {   // example3::test_feasible — bsearch over BSearchAdaptor (clone + set_xc on the device handle)
        Options o; o.tolerance = 1e-8;
        BSearchAdaptor<Example3, EllCuda> adaptor{Example3{}, EllCuda::new_with_scalar(100.0, {0.0, 0.0}), o};
        auto [feasible, num_iters] = bsearch(adaptor, {-100.0, 100.0}, o);
        EXPECT(feasible);
        EXPECT(num_iters == 34);
    }
    {   // cutting_plane_tests::test_cutting_plane_feas_no_soln
        auto ellip = EllCuda::new_with_scalar(10.0, {0.0, 0.0});
        AlwaysCut omega;
        Options o; o.max_iters = 200;
        auto [x, niter] = cutting_plane_feas(omega, ellip, o);
        EXPECT(!x.has_value());
        EXPECT(niter == 2);
    }
    {   // ell.rs::test_update_central_cut (exact assert_eq! on the matrix)
        auto ellip = EllCuda::new_with_scalar(0.01, Arr(4, 0.0));
        const CutStatus st = ellip.update_central_cut(std::make_pair(Arr(4, 0.5), SingleCut{0.0}));
        EXPECT(st == CutStatus::Success);
        const Arr m = ellip.mq(), x = ellip.xc();
        for (int i = 0; i < 4; ++i) {
            EXPECT(x[i] == -0.01 * 1.0);
            for (int j = 0; j < 4; ++j) EXPECT(m[i * 4 + j] == (i == j ? 1.0 : 0.0) - 0.1 * 1.0);
        }
        bool threw = false;
        try { ellip.update_bias_cut(std::make_pair(Arr(5, 0.5), SingleCut{0.0})); } catch (const std::invalid_argument&) { threw = true; }
        EXPECT(threw);  // the reference panics on a shape mismatch (src/arr.rs:427-429)
    }
    {   // quasicvx::test_feasible (Ell, 35 iterations) and ::test_feasible_stable (EllStable), src/quasicvx.rs:61-105
        struct Quasicvx {
            int idx = -1;
            std::pair<std::pair<Arr, SingleCut>, bool> assess_optim(const Arr& xc, double& gamma) {
                const double sqrtx = xc[0], logy = xc[1];
                for (int k = 0; k < 2; ++k) {
                    idx = idx == 1 ? 0 : idx + 1;
                    const double fv = idx == 0 ? sqrtx * sqrtx - logy : -sqrtx + gamma * std::exp(logy);
                    if (fv > 0.0) {
                        const Arr g = idx == 0 ? Arr{2.0 * sqrtx, -1.0} : Arr{-1.0, gamma * std::exp(logy)};
                        return {{g, SingleCut{fv}}, false};
                    }
                }
                gamma = sqrtx / std::exp(logy);
                return {{{-1.0, sqrtx}, SingleCut{0.0}}, true};
            }
        };
        Options o; o.tolerance = 1e-8;
        auto ell = EllCuda::new_({10.0, 10.0}, {0.0, 0.0});
        Quasicvx q1;
        double gamma = 0.0;
        auto [x, niter] = cutting_plane_optim(q1, ell, gamma, o);
        EXPECT(x.has_value() && niter == 35);
        EXPECT((*x)[0] * (*x)[0] >= 0.49 && (*x)[0] * (*x)[0] <= 0.51 && std::exp((*x)[1]) >= 1.6 && std::exp((*x)[1]) <= 1.7);
        auto stable = EllStableCuda::new_({10.0, 10.0}, {0.0, 0.0});
        Quasicvx q2;
        double gamma2 = 0.0;
        auto [xs, niter_s] = cutting_plane_optim(q2, stable, gamma2, o);
        EXPECT(xs.has_value());
        EXPECT(niter_s == 49);   // not pinned by the reference; value of the CPU restatement
        auto inf = EllStableCuda::new_with_scalar(10.0, {100.0, 100.0});
        Quasicvx q3;
        double gamma3 = 0.0;
        EXPECT(!cutting_plane_optim(q3, inf, gamma3, Options{}).first.has_value());
    }
    std::printf("host mirror: all reference driver tests passed on the GPU search space\n");
    return 0;
}
