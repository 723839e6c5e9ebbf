// ellalgo_cuda.hpp — C++ host mirror of the reference crate's plugin surface over the C ABI of ellcuda.h.
//
// The reference is Rust; this image has no Rust toolchain, so the host side above the C ABI is C++ with the same
// names, argument meaning and error behaviour as the Rust items it mirrors (paths relative to the reference root):
//
//   SingleCut / ParallelCut / CutStatus / Options          src/cutting_plane.rs:9, :18, :32-37, :50-100
//   SearchSpace methods on EllCuda                          src/cutting_plane.rs:154-182, src/ell.rs:140-229
//   Ell::new / new_with_scalar / new_with_matrix / from_covariance   src/ell.rs:31-78
//   cutting_plane_feas / _optim / _optim_q                  src/cutting_plane.rs:205-227, 286-313, 331-374
//   BSearchAdaptor / bsearch                                src/cutting_plane.rs:376-466
//
// An oracle is any type with the reference's methods:
//   std::optional<std::pair<Arr, Cut>> assess_feas(const Arr& xc);                (OracleFeas,  :119-126; optional update(gamma))
//   std::pair<std::pair<Arr, Cut>, bool> assess_optim(const Arr& xc, double& gamma);   (OracleOptim, :129-136)
//   QAnswer<Cut> assess_optim_q(const Arr& xc, double& gamma, bool retry);        (OracleOptimQ, :139-147)
// where Cut is SingleCut or ParallelCut.  Shape errors panic in Rust (src/arr.rs:427-429); here they throw
// std::invalid_argument, and infrastructure errors (no GPU, CUDA failure) throw ellalgo::CudaError — there is no
// CPU fallback.  Header-only; link with libellcuda.so.
#pragma once
#include <cstddef>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "ellcuda.h"

namespace ellalgo {

using Arr = std::vector<double>;

struct SingleCut {  // src/cutting_plane.rs:9
    double beta;
};
struct ParallelCut {  // src/cutting_plane.rs:18
    double beta0;
    std::optional<double> beta1;
};
enum class CutStatus : int { Success = 0, NoSoln = 1, NoEffect = 2, Unknown = 3 };  // :32-37
struct Options {  // :50-100
    std::size_t max_iters = 2000;
    double tolerance = 1e-20;
    bool verbose = false;
};

struct CudaError : std::runtime_error {
    int code;
    CudaError(int c, const std::string& m) : std::runtime_error("libellcuda error " + std::to_string(c) + ": " + m), code(c) {}
};
inline void check(int rc) {
    if (rc == ELLCUDA_E_ARG) throw std::invalid_argument(ellcuda_last_error());
    if (rc != ELLCUDA_OK) throw CudaError(rc, ellcuda_last_error());
}

namespace detail {
struct CutArgs { double b0; int has1; double b1; };
inline CutArgs cut_args(const SingleCut& c) { return {c.beta, -1, 0.0}; }
inline CutArgs cut_args(const ParallelCut& c) { return {c.beta0, c.beta1 ? 1 : 0, c.beta1.value_or(0.0)}; }
}  // namespace detail

struct EllConfig {
    int device = 0;
    bool no_defer_trick = false;    // src/ell.rs:10
    bool use_parallel_cut = true;   // src/ell_calc.rs:630
    int reduce_mode = 0;            // 0 = bit-identical to the reference's summation order
    int storage = ELL_STORAGE_FULL; // ELL_STORAGE_SYMMETRIC: lower block-triangle only, half the HBM traffic, results to 1e-9
    std::vector<int> devices;       // non-empty: rows sharded over these devices of this process (ellcuda_create_sharded)
};

// The reference's `Ell` (KIND = ELL_KIND_ELL, src/ell.rs) with its matrix in B200 HBM, or `EllStable`
// (KIND = ELL_KIND_ELLSTABLE, src/ell_stable.rs, n <= 2048).  Copying clones the device state (#[derive(Clone)]).
template <int KIND>
class EllCudaT {
  public:
    using EllCuda = EllCudaT;   // the constructors below are written once for both kinds
    using Config = EllConfig;
    static EllCuda new_with_matrix(double kappa, const Arr& mq, const Arr& xc, Config cfg = {}) {  // src/ell.rs:31-41
        if (mq.size() != xc.size() * xc.size()) throw std::invalid_argument("mq must be n x n");
        return EllCuda(kappa, &mq, nullptr, xc, cfg);
    }
    static EllCuda new_(const Arr& val, const Arr& xc, Config cfg = {}) {  // Ell::new, src/ell.rs:54-56
        if (val.size() != xc.size()) throw std::invalid_argument("val must have n entries");
        return EllCuda(1.0, nullptr, &val, xc, cfg);
    }
    static EllCuda new_with_scalar(double val, const Arr& xc, Config cfg = {}) {  // src/ell.rs:65-67
        return EllCuda(val, nullptr, nullptr, xc, cfg);
    }
    static EllCuda from_covariance(const Arr& cov, const Arr& xc, Config cfg = {}) {  // src/ell.rs:76-78
        return new_with_matrix(1.0, cov, xc, cfg);
    }
    EllCudaT(const EllCudaT& o) : n_(o.n_) { check(ellcuda_clone(o.h_, &h_)); }
    EllCudaT(EllCudaT&& o) noexcept : h_(o.h_), n_(o.n_) { o.h_ = nullptr; }
    EllCuda& operator=(EllCuda o) noexcept { std::swap(h_, o.h_); std::swap(n_, o.n_); return *this; }
    ~EllCudaT() { if (h_) ellcuda_destroy(h_); }

    // ---- SearchSpace (src/cutting_plane.rs:154-182)
    Arr xc() const { Arr x(n_); check(ellcuda_xc(h_, x.data())); return x; }
    double tsq() const { double v; check(ellcuda_tsq(h_, &v)); return v; }
    template <class Cut> CutStatus update_bias_cut(const std::pair<Arr, Cut>& cut) { return update(ELL_UPDATE_BIAS, cut); }
    template <class Cut> CutStatus update_central_cut(const std::pair<Arr, Cut>& cut) { return update(ELL_UPDATE_CENTRAL, cut); }
    template <class Cut> CutStatus update_q(const std::pair<Arr, Cut>& cut) { return update(ELL_UPDATE_Q, cut); }
    void set_xc(const Arr& x) {
        if (x.size() != n_) throw std::invalid_argument("xc must have n entries");
        check(ellcuda_set_xc(h_, x.data()));
    }
    // ---- pub fields of Ell (src/ell.rs:9-16)
    double kappa() const { double v; check(ellcuda_kappa(h_, &v)); return v; }
    Arr mq() const { Arr m(n_ * n_); check(ellcuda_get_mq(h_, m.data())); return m; }
    // `pub no_defer_trick` is assigned after construction in the reference (src/ell.rs:342-344); so is the helper's
    // `use_parallel_cut` (src/ell_calc.rs:630)
    void set_no_defer_trick(bool on) { check(ellcuda_set_no_defer_trick(h_, on)); }
    void set_use_parallel_cut(bool on) { check(ellcuda_set_use_parallel_cut(h_, on)); }
    std::size_t len() const { return n_; }
    ell_handle handle() const { return h_; }

  private:
    EllCudaT(double kappa, const Arr* mq, const Arr* diag, const Arr& xc, Config cfg) : n_(xc.size()) {
        ell_create_desc d{};
        d.kind = KIND; d.device = cfg.device; d.n = static_cast<int64_t>(n_); d.kappa = kappa;
        d.mq = mq ? mq->data() : nullptr; d.diag = diag ? diag->data() : nullptr; d.xc = xc.data();
        d.no_defer_trick = cfg.no_defer_trick; d.use_parallel_cut = cfg.use_parallel_cut;
        d.row_begin = 0; d.row_end = static_cast<int64_t>(n_); d.reduce_mode = cfg.reduce_mode; d.storage = cfg.storage; d.stream = nullptr;
        if (!cfg.devices.empty()) {
            std::vector<int32_t> devs(cfg.devices.begin(), cfg.devices.end());
            check(ellcuda_create_sharded(&d, static_cast<int32_t>(devs.size()), devs.data(), &h_));
        } else {
            check(ellcuda_create(&d, &h_));
        }
    }
    template <class Cut> CutStatus update(int which, const std::pair<Arr, Cut>& cut) {
        if (cut.first.size() != n_) throw std::invalid_argument("grad must have n entries");  // assert_eq!(self.cols, x.len())
        const detail::CutArgs a = detail::cut_args(cut.second);
        int32_t st = ELL_UNKNOWN;
        check(ellcuda_update(h_, which, cut.first.data(), a.b0, a.has1, a.b1, &st));
        return static_cast<CutStatus>(st);
    }
    ell_handle h_ = nullptr;
    std::size_t n_ = 0;
};
using EllCuda = EllCudaT<ELL_KIND_ELL>;
using EllStableCuda = EllCudaT<ELL_KIND_ELLSTABLE>;   // from_covariance is Ell's only; mq() returns the factor

// ---- drivers (generic over any Space with the SearchSpace methods, any Oracle with the trait's method)
template <class Oracle, class Space>
std::pair<std::optional<Arr>, std::size_t> cutting_plane_feas(Oracle& omega, Space& space, const Options& options) {  // :205-227
    for (std::size_t niter = 0; niter < options.max_iters; ++niter) {
        auto cut = omega.assess_feas(space.xc());
        if (!cut) return {space.xc(), niter};
        const CutStatus status = space.update_bias_cut(*cut);
        if (status != CutStatus::Success || space.tsq() < options.tolerance) return {std::nullopt, niter};
    }
    return {std::nullopt, options.max_iters};
}

template <class Oracle, class Space>
std::pair<std::optional<Arr>, std::size_t> cutting_plane_optim(Oracle& omega, Space& space, double& gamma,
                                                                const Options& options) {  // :286-313
    std::optional<Arr> x_best;
    for (std::size_t niter = 0; niter < options.max_iters; ++niter) {
        auto [cut, shrunk] = omega.assess_optim(space.xc(), gamma);
        CutStatus status;
        if (shrunk) {
            x_best = space.xc();
            status = space.update_central_cut(cut);
        } else {
            status = space.update_bias_cut(cut);
        }
        if (status != CutStatus::Success || space.tsq() < options.tolerance) return {x_best, niter};
    }
    return {x_best, options.max_iters};
}

template <class Cut>
struct QAnswer {  // the 4-tuple of OracleOptimQ::assess_optim_q (:139-147)
    std::pair<Arr, Cut> cut;
    bool shrunk;
    Arr x_q;
    bool more_alt;
};

template <class Oracle, class Space>
std::pair<std::optional<Arr>, std::size_t> cutting_plane_optim_q(Oracle& omega, Space& space_q, double& gamma,
                                                                  const Options& options) {  // :331-374
    std::optional<Arr> x_best;
    bool retry = false;
    for (std::size_t niter = 0; niter < options.max_iters; ++niter) {
        auto a = omega.assess_optim_q(space_q.xc(), gamma, retry);
        if (a.shrunk) {
            x_best = a.x_q;
            retry = false;
        }
        const CutStatus status = space_q.update_q(a.cut);
        if (status == CutStatus::Success) {
            retry = false;
        } else if (status == CutStatus::NoSoln) {
            return {x_best, niter};
        } else if (status == CutStatus::NoEffect) {
            if (!a.more_alt) return {x_best, niter};
            retry = true;
        }
        if (space_q.tsq() < options.tolerance) return {x_best, niter};
    }
    return {x_best, options.max_iters};
}

template <class Oracle, class Space>
struct BSearchAdaptor {  // :376-419 — Space must be copyable (Clone)
    Oracle omega;
    Space space;
    Options options;
    bool assess_bs(double gamma) {
        Space sp = space;
        omega.update(gamma);
        auto r = cutting_plane_feas(omega, sp, options);
        if (r.first) {
            space.set_xc(*r.first);
            return true;
        }
        return false;
    }
};

template <class OracleBS>
std::pair<bool, std::size_t> bsearch(OracleBS& omega, std::pair<double, double> intrvl, const Options& options) {  // :441-466
    double lower = intrvl.first, upper = intrvl.second;
    if (!(lower <= upper)) throw std::invalid_argument("assertion failed: lower <= upper");
    const double u_orig = upper;
    for (std::size_t niter = 0; niter < options.max_iters; ++niter) {
        const double tau = (upper - lower) / 2.0;
        if (tau < options.tolerance) return {upper != u_orig, niter};
        double gamma = lower;
        gamma += tau;
        if (omega.assess_bs(gamma)) upper = gamma; else lower = gamma;
    }
    return {upper != u_orig, options.max_iters};
}

}  // namespace ellalgo
