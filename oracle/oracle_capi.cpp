// oracle_capi.cpp — extern "C" face of the CPU restatement, loaded by tests/ and bench.py via ctypes.
// TEST INFRASTRUCTURE, NOT PRODUCT (see ellref.hpp).  Uses include/ellcuda.h only for the shared
// plain-C descriptor structs and constants, so that the same inputs drive the checker and the product.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <memory>
#include <thread>

#include "../include/ellcuda.h"
#include "ellref_linalg.hpp"
#include "ellref_oracles.hpp"

using namespace ellref;

namespace {

inline const double* at(const double* p, int64_t stride, int64_t i) { return p ? p + stride * i : nullptr; }

template <class Space>
Space make_space(const ell_batch_desc& d, int64_t i) {
    const std::size_t n = static_cast<std::size_t>(d.n);
    const double kappa = *at(d.kappa0, d.kappa0_stride, i);
    const double* x0 = at(d.xc0, d.xc0_stride, i);
    Vec xc(x0, x0 + n);
    Vec m(n * n, 0.0);
    if (d.mq0) {
        const double* m0 = at(d.mq0, d.mq0_stride, i);
        std::copy(m0, m0 + n * n, m.begin());
    } else if (d.diag0) {
        const double* dg = at(d.diag0, d.diag0_stride, i);
        for (std::size_t k = 0; k < n; ++k) m[k * n + k] = dg[k];
    } else {
        for (std::size_t k = 0; k < n; ++k) m[k * n + k] = 1.0;
    }
    Space sp(kappa, std::move(m), std::move(xc));
    sp.set_use_parallel_cut(d.use_parallel_cut != 0);
    if constexpr (std::is_same_v<Space, Ell>) sp.no_defer_trick = d.no_defer_trick != 0;
    return sp;
}

template <class Space>
void store(const ell_batch_desc& d, const ell_batch_result& r, int64_t i, const RunResult& rr, const Space& sp, double gamma) {
    const std::size_t n = static_cast<std::size_t>(d.n);
    if (r.has_xbest) r.has_xbest[i] = rr.xbest ? 1 : 0;
    if (r.xbest && rr.xbest) std::copy(rr.xbest->begin(), rr.xbest->end(), r.xbest + i * n);
    if (r.num_iters) r.num_iters[i] = static_cast<int64_t>(rr.num_iters);
    if (r.exit_status) r.exit_status[i] = rr.exit_status;
    if (r.gamma) r.gamma[i] = gamma;
    if (r.tsq) r.tsq[i] = sp.tsq;
    if (r.xc) std::copy(sp.xc.begin(), sp.xc.end(), r.xc + i * n);
    if (r.kappa) r.kappa[i] = sp.kappa;
}

template <class Space>
int solve_one(const ell_batch_desc& d, const ell_batch_result& r, int64_t i) {
    Space sp = make_space<Space>(d, i);
    const Options opt{static_cast<std::size_t>(d.max_iters), d.tolerance};
    double gamma = (d.gamma0 && d.mode != ELL_MODE_FEAS) ? *at(d.gamma0, d.gamma0_stride, i) : 0.0;
    const double* p = at(d.params, d.params_stride, i);
    const std::size_t n = static_cast<std::size_t>(d.n);
    RunResult rr;
    auto optim = [&](auto&& omega) { rr = cutting_plane_optim(omega, sp, gamma, opt); };
    auto feas = [&](auto&& omega) { rr = cutting_plane_feas(omega, sp, opt); };
    auto optq = [&](auto&& omega) { rr = cutting_plane_optim_q(omega, sp, gamma, opt); };
    switch (d.oracle_id) {
        case ELL_ORACLE_EXAMPLE1: { Example1 o; optim(o); break; }
        case ELL_ORACLE_EXAMPLE1_RR: { Example1RR o; optim(o); break; }
        case ELL_ORACLE_EXAMPLE4: { Example4 o; optim(o); break; }
        case ELL_ORACLE_QUASICVX: { Quasicvx o; optim(o); break; }
        case ELL_ORACLE_PROFIT: { Profit o(p[0], p[1], p[2], p + 3, p + 5); optim(o); break; }
        case ELL_ORACLE_PROFIT_RB: { ProfitRb o(p[0], p[1], p[2], p + 3, p + 5, p + 7); optim(o); break; }
        case ELL_ORACLE_PROFIT_Q: { ProfitQ o(p[0], p[1], p[2], p + 3, p + 5); optq(o); break; }
        case ELL_ORACLE_LOWPASS: {
            if (d.aux[0] != 15 * d.n) return ELLCUDA_E_ARG;
            Lowpass o(n, d.table, p[0], p[1], p[2], p[3], p[4]);
            optim(o);
            break;
        }
        case ELL_ORACLE_BOWL: { Bowl o{p, n}; optim(o); break; }
        case ELL_ORACLE_LMI: {
            const std::size_t m = static_cast<std::size_t>(d.aux[0]);
            LMI o(d.table, d.table + n * m * m, n, m);
            feas(o);
            break;
        }
        case ELL_ORACLE_EXAMPLE2: { Example2 o; feas(o); break; }
        case ELL_ORACLE_EXAMPLE3: {
            if (d.mode == ELL_MODE_BSEARCH) {   // src/example3.rs:68-85 generalised: interval [gamma0, p[0]], outer Options{p[2], p[1]}
                BSearchAdaptor<Example3, Space> ad{Example3{}, sp, opt};
                double upper = p[0];
                const Options outer{static_cast<std::size_t>(p[2]), p[1]};
                auto r = bsearch(ad, {gamma, p[0]}, outer, &upper);
                sp = ad.space;
                rr.num_iters = r.second;
                rr.exit_status = r.second < outer.max_iters ? ExitTol : ExitMaxIt;
                if (r.first) rr.xbest = sp.xc;
                gamma = upper;
                break;
            }
            Example3 o; if (p) o.target = p[0]; feas(o); break;
        }
        case ELL_ORACLE_HALFPLANE: { HalfPlane o; feas(o); break; }
        case ELL_ORACLE_ALWAYS_CUT: {
            AlwaysCut o;
            if (d.mode == ELL_MODE_FEAS) feas(o); else if (d.mode == ELL_MODE_OPTIM_Q) optq(o); else optim(o);
            break;
        }
        case ELL_ORACLE_OPTIM_BOX: {
            if (d.mode == ELL_MODE_OPTIM_Q) { OptimBoxQ o; optq(o); } else { OptimBox o; optim(o); }
            break;
        }
        case ELL_ORACLE_LMI_LINOPT: {
            auto run = [&](auto o) {
                o.c.assign(d.table, d.table + n);
                const double* t = d.table + n;
                for (int b = 0; b < d.aux[0]; ++b) {
                    const std::size_t m = static_cast<std::size_t>(d.aux[1 + b]);
                    o.blocks.emplace_back(t, t + n * m * m, n, m);
                    t += (n + 1) * m * m;
                }
                optim(o);
            };
            // aux[7] = 1 (checker only): the blocks are LMIOldOracle instead of LMIOracle (tests/lmi_tests.rs:208-225)
            if (d.aux[7] == 1) run(LmiLinOptT<LMIOld>{}); else run(LmiLinOpt{});
            break;
        }
        case ELL_ORACLE_LMI_OLD: {
            const std::size_t m = static_cast<std::size_t>(d.aux[0]);
            LMIOld o(d.table, d.table + n * m * m, n, m);
            feas(o);
            break;
        }
        case ELL_ORACLE_LMI0: {
            const std::size_t m = static_cast<std::size_t>(d.aux[0]);
            LMI0 o(d.table, nullptr, n, m);
            feas(o);
            break;
        }
        default: return ELLCUDA_E_ARG;
    }
    store(d, r, i, rr, sp, gamma);
    return 0;
}

}  // namespace

extern "C" {

// Same contract as ellcuda_batch_solve (include/ellcuda.h) on host pointers, `nthreads` std::threads
// over instances (0 ⇒ hardware_concurrency).  Returns the number of threads used, or <0.
int ellref_batch_solve(const ell_batch_desc* d, const ell_batch_result* r, int nthreads) {
    if (!d || !r || d->n <= 0 || d->count < 0) return ELLCUDA_E_ARG;
    int T = nthreads > 0 ? nthreads : static_cast<int>(std::thread::hardware_concurrency());
    if (T < 1) T = 1;
    if (static_cast<int64_t>(T) > d->count) T = static_cast<int>(std::max<int64_t>(1, d->count));
    std::atomic<int64_t> next{0};
    std::atomic<int> err{0};
    auto work = [&]() {
        for (;;) {
            const int64_t i0 = next.fetch_add(16);
            if (i0 >= d->count) break;
            const int64_t i1 = std::min<int64_t>(d->count, i0 + 16);
            for (int64_t i = i0; i < i1; ++i) {
                const int rc = d->kind == ELL_KIND_ELLSTABLE ? solve_one<EllStable>(*d, *r, i) : solve_one<Ell>(*d, *r, i);
                if (rc) err = rc;
            }
        }
    };
    if (T == 1) {
        work();
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back(work);
        for (auto& t : th) t.join();
    }
    return err ? err.load() : T;
}

int ellref_hardware_threads() { return static_cast<int>(std::thread::hardware_concurrency()); }

// ---- single search space, for step-by-step state parity ---------------------------------------------
struct ellref_space {
    int kind;
    std::unique_ptr<Ell> ell;
    std::unique_ptr<EllStable> stable;
};

ellref_space* ellref_space_new(int kind, int64_t n, double kappa, const double* mq, const double* diag, const double* xc,
                               int no_defer_trick, int use_parallel_cut) {
    ell_batch_desc d{};
    d.kind = kind; d.n = static_cast<int32_t>(n); d.count = 1;
    d.kappa0 = &kappa; d.xc0 = xc; d.diag0 = diag; d.mq0 = mq;
    d.no_defer_trick = no_defer_trick; d.use_parallel_cut = use_parallel_cut;
    auto* s = new ellref_space{kind, nullptr, nullptr};
    if (kind == ELL_KIND_ELLSTABLE) s->stable = std::make_unique<EllStable>(make_space<EllStable>(d, 0));
    else s->ell = std::make_unique<Ell>(make_space<Ell>(d, 0));
    return s;
}
void ellref_space_free(ellref_space* s) { delete s; }
ellref_space* ellref_space_clone(const ellref_space* s) {
    auto* c = new ellref_space{s->kind, nullptr, nullptr};
    if (s->ell) c->ell = std::make_unique<Ell>(*s->ell);
    if (s->stable) c->stable = std::make_unique<EllStable>(*s->stable);
    return c;
}
int ellref_space_update(ellref_space* s, int which, const double* grad, double beta0, int has_beta1, double beta1) {
    const std::size_t n = s->ell ? s->ell->n : s->stable->n;
    Vec g(grad, grad + n);
    const Cut c = has_beta1 > 0 ? Cut::parallel(beta0, beta1) : Cut::single(beta0);
    return s->ell ? s->ell->update(which, g, c) : s->stable->update(which, g, c);
}
void ellref_space_get(const ellref_space* s, double* xc, double* mq, double* kappa, double* tsq) {
    const Vec& x = s->ell ? s->ell->xc : s->stable->xc;
    const Vec& m = s->ell ? s->ell->mq : s->stable->mq;
    if (xc) std::copy(x.begin(), x.end(), xc);
    if (mq) std::copy(m.begin(), m.end(), mq);
    if (kappa) *kappa = s->ell ? s->ell->kappa : s->stable->kappa;
    if (tsq) *tsq = s->ell ? s->ell->tsq : s->stable->tsq;
}
void ellref_space_set_xc(ellref_space* s, const double* xc) {
    Vec& x = s->ell ? s->ell->xc : s->stable->xc;
    std::copy(xc, xc + x.size(), x.begin());
}

// The reference's large-n update restated once more with the matrix held by the caller (so a 2 GiB M is
// not copied): exactly Ell::update_core (src/ell.rs:97-137) on (mq, xc, kappa, tsq) in place.
int ellref_ell_update_inplace(int64_t n_, double* mq, double* xc, double* kappa, double* tsq, int no_defer_trick,
                              int use_parallel_cut, int which, const double* grad, double beta0, int has_beta1, double beta1,
                              double* grad_t_out) {
    const std::size_t n = static_cast<std::size_t>(n_);
    EllCalc helper(n);
    helper.use_parallel_cut = use_parallel_cut != 0;
    Vec grad_t(n);
    dot_mv(mq, grad, n, n, grad_t.data());
    if (grad_t_out) std::copy(grad_t.begin(), grad_t.end(), grad_t_out);
    const double omega = dot(grad, grad_t.data(), n);
    *tsq = *kappa * omega;
    const Cut c = has_beta1 > 0 ? Cut::parallel(beta0, beta1) : Cut::single(beta0);
    const CalcResult r = helper.dispatch(which, c, *tsq);
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
    *kappa *= r.v.delta;
    if (no_defer_trick) {
        for (std::size_t k = 0; k < n * n; ++k) mq[k] *= *kappa;
        *kappa = 1.0;
    }
    return r.status;
}

// ---- EllCalc known-answer access: fn 0 bias, 1 central, 2 q (Cut dispatch), 3 core.parallel_bias_cut_fast,
//      4 core.bias_cut_fast, 5 constants (cst1, cst2, half_n) ----------------------------------------------
int ellref_calc(int64_t n, int use_parallel_cut, int fn, double b0, int has_b1, double b1, double tsq, double a3, double a4,
                double* out3) {
    EllCalc h(static_cast<std::size_t>(n));
    h.use_parallel_cut = use_parallel_cut != 0;
    CalcResult r{Success, {0, 0, 0}};
    if (fn <= 2) {
        const Cut c = has_b1 > 0 ? Cut::parallel(b0, b1) : Cut::single(b0);
        r = h.dispatch(fn, c, tsq);
    } else if (fn == 3) {
        r.v = h.helper.calc_parallel_bias_cut_fast(b0, b1, tsq, a3, a4);
    } else if (fn == 4) {
        r.v = h.helper.calc_bias_cut_fast(b0, b1, tsq);  // (beta, tau, eta)
    } else {
        r.v = {h.helper.cst1, h.helper.cst2, h.helper.half_n};
    }
    out3[0] = r.v.rho; out3[1] = r.v.sigma; out3[2] = r.v.delta;
    return r.status;
}

// ---- LDLTMgr known-answer access ---------------------------------------------------------------------
// returns is_spd; pos[2]; if not spd and want_witness: wit[n], *ep
int ellref_ldlt(int64_t n_, const double* mat, int allow_semidefinite, int64_t* pos, double* wit, double* ep) {
    const std::size_t n = static_cast<std::size_t>(n_);
    LDLTMgr mgr(n);
    auto get = [&](std::size_t i, std::size_t j) { return mat[i * n + j]; };
    const bool spd = allow_semidefinite ? mgr.factor_with_allow_semidefinite(get) : mgr.factor(get);
    pos[0] = static_cast<int64_t>(mgr.pos0);
    pos[1] = static_cast<int64_t>(mgr.pos1);
    if (!spd) {
        const double e = mgr.witness();
        if (ep) *ep = e;
        if (wit) std::copy(mgr.wit.begin(), mgr.wit.end(), wit);
    }
    return spd ? 1 : 0;
}

// {LMIOracle, LMIOldOracle, LMI0Oracle}::assess_feas once (variant 0, 1, 2; b is unused for LMI0); returns 1 if a cut was
// produced (g[n], *beta), 0 if feasible
int ellref_lmi_assess_variant(int variant, int64_t nvar, int64_t m, const double* f, const double* b, const double* xc, double* g,
                              double* beta) {
    Vec x(xc, xc + nvar);
    FeasAnswer c;
    const std::size_t nv = static_cast<std::size_t>(nvar), mm = static_cast<std::size_t>(m);
    if (variant == 1) { LMIOld o(f, b, nv, mm); c = o.assess_feas(x); }
    else if (variant == 2) { LMI0 o(f, b, nv, mm); c = o.assess_feas(x); }
    else { LMI o(f, b, nv, mm); c = o.assess_feas(x); }
    if (!c) return 0;
    std::copy(c->first.begin(), c->first.end(), g);
    *beta = c->second.beta0;
    return 1;
}
int ellref_lmi_assess(int64_t nvar, int64_t m, const double* f, const double* b, const double* xc, double* g, double* beta) {
    return ellref_lmi_assess_variant(0, nvar, m, f, b, xc, g, beta);
}

// ProfitOracle::assess_optim once on a fresh oracle (profit_oracle.rs:262-277): returns shrunk
int ellref_profit_assess(const double* p7, const double* y, double* gamma, double* g2, double* beta) {
    Profit o(p7[0], p7[1], p7[2], p7 + 3, p7 + 5);
    Vec yy(y, y + 2);
    auto a = o.assess_optim(yy, *gamma);
    g2[0] = a.first.first[0]; g2[1] = a.first.first[1];
    *beta = a.first.second.beta0;
    return a.second ? 1 : 0;
}

// the host libm the restatement itself uses (what Rust's f64::exp / f64::ln call): which = 0 exp, 1 ln
void ellref_libm(int which, const double* x, double* y, int64_t n) {
    for (int64_t i = 0; i < n; ++i) y[i] = which == 0 ? std::exp(x[i]) : std::log(x[i]);
}

// power_iteration variants 1..4 (src/power_iteration.rs): x in/out, returns niter, *eigenvalue
int64_t ellref_power_iteration(int variant, int64_t n, const double* a, double* x, int64_t max_iters, double tol, double* eig) {
    Vec A(a, a + n * n), X(x, x + n);
    const Options o{static_cast<std::size_t>(max_iters), tol};
    std::pair<double, std::size_t> r;
    switch (variant) {
        case 1: r = power_iteration(A, X, o); break;
        case 2: r = power_iteration2(A, X, o); break;
        case 3: r = power_iteration3(A, X, o); break;
        default: r = power_iteration4(A, X, o); break;
    }
    std::copy(X.begin(), X.end(), x);
    *eig = r.first;
    return static_cast<int64_t>(r.second);
}
// conjugate_gradient (src/conjugate_gradient.rs): x in (x0) / out; returns iterations, -1 / -2 for the two Err cases
int64_t ellref_conjugate_gradient(int64_t n, const double* a, const double* b, double* x, double tol, int64_t max_iter) {
    Vec A(a, a + n * n), B(b, b + n), X(x, x + n);
    const long r = conjugate_gradient(A, B, X, tol, static_cast<std::size_t>(max_iter));
    std::copy(X.begin(), X.end(), x);
    return r;
}

void ellref_lowpass_table(int64_t n, double* out) {
    const Vec s = lowpass_spectrum(static_cast<std::size_t>(n));
    std::copy(s.begin(), s.end(), out);
}
// Arr::dot (arr.rs:443-451), Arr::dot_mv (:426-442), linspace (:506-520) on their own (arr.rs:676-709 known answers)
double ellref_dot(const double* a, const double* b, int64_t n) { return dot(a, b, static_cast<std::size_t>(n)); }
void ellref_dot_mv(const double* m, const double* x, int64_t rows, int64_t cols, double* out) {
    dot_mv(m, x, static_cast<std::size_t>(rows), static_cast<std::size_t>(cols), out);
}
void ellref_linspace(double start, double end, int64_t n, double* out) {
    const Vec v = linspace(start, end, static_cast<std::size_t>(n));
    std::copy(v.begin(), v.end(), out);
}
int ellref_lowpass_band_edge(int64_t n, double w) { return lowpass_band_edge(static_cast<std::size_t>(n), w); }
void ellref_lowpass_case_params(double* out5) { lowpass_case_params(out5); }

// bsearch over BSearchAdaptor<Example3, Ell> (src/example3.rs:68-85): returns niter, *feasible
int64_t ellref_bsearch_example3(double kappa, const double* xc2, double lo, double hi, double tol_inner, int64_t max_inner,
                                double tol_outer, int64_t max_outer, int* feasible) {
    BSearchAdaptor<Example3, Ell> ad{Example3{}, Ell::with_scalar(kappa, Vec{xc2[0], xc2[1]}),
                                     Options{static_cast<std::size_t>(max_inner), tol_inner}};
    auto r = bsearch(ad, {lo, hi}, Options{static_cast<std::size_t>(max_outer), tol_outer});
    *feasible = r.first ? 1 : 0;
    return static_cast<int64_t>(r.second);
}
int64_t ellref_bsearch_bs2(double lo, double hi, double tol, int64_t max_iters, int* feasible) {
    BS2 o;
    auto r = bsearch(o, {lo, hi}, Options{static_cast<std::size_t>(max_iters), tol});
    *feasible = r.first ? 1 : 0;
    return static_cast<int64_t>(r.second);
}

}  // extern "C"
