// mpc_b200.hpp — C++17 host-side mirror of the reference's public types over the C ABI of mpc_b200.h.
//
// The reference is a compiled (Rust) crate; no Rust toolchain exists in this image or on the GPU box, so the host
// side above the C ABI that is BUILT and TESTED is this header (the Rust shim under rust/ is the same thing as
// source only).  Names, argument order and error behaviour follow the reference:
//
//   reference                                                       here
//   mpc::mppi::Mppi::<N,K,S>::new(dynamics, cost, λ, σ, limit)        mpc::mppi::Mppi<N,K,S>::create(...)      src/mppi.rs:16-30
//   mppi.compute(&x, &u_n) -> Result<SVector<f64,N>, &'static str>    mppi.compute(x, u_n) -> Result<VecN>      src/mppi.rs:33-92
//   mpc::ukf::UnscentedKalmanFilter::new(x, p, q, r)   (n=4, o=3)     mpc::ukf::UnscentedKalmanFilter::create   src/ukf.rs:30-42
//   ukf.predict(u, fx) / update(&z, hx) / state() / covariance()      same names                               src/ukf.rs:44-94
//   mpc::ukf2::UnscentedKalmanFilter (n=6, o=5) + set_q               + set_q, set_r, set_enable               src/ukf2.rs:30-98
//   mpc::gaussian::Gaussian (+, -, *, * f64, Default)                  mpc::gaussian::Gaussian                  src/gaussian.rs:1-63
//
// `new` is a C++ keyword, hence `create`.  The reference passes `dynamics`/`cost`/`fx`/`hx` as host fn pointers or
// closures; a kernel cannot call those, so they are DeviceModel tags naming the built-in device models (SURVEY.md
// appendix A).  Vectors are std::array<double, N>; matrices are row-major std::array<double, R*C> (element (i,j) at
// i*C + j).  Result<T> plays Rust's Result<T, &'static str>: `ok`, `value`, `err` (the reference's message), and
// `unwrap()` throws std::runtime_error(err) like `.unwrap()` panics.  UKF failures that panic in the reference
// ("Inverse fail", src/ukf.rs:69) throw std::runtime_error with the same message.
// There is no CPU fallback: construction throws when libmpc_b200 finds no CUDA device.
#ifndef MPC_B200_HPP
#define MPC_B200_HPP

#include <array>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "mpc_b200.h"

namespace mpc {

enum class DeviceModel : int32_t {
    L = MPCB_MODEL_L,              // examples/mppi4.rs:20-27,73-89
    NL = MPCB_MODEL_NL,            // examples/mppi4-non-liner.rs:20-27,73-94
    NL6 = MPCB_MODEL_NL6,          // examples/mppi4-non-liner-ukf.rs:33-35,126-148
    PEN_LIN = MPCB_MODEL_PEN_LIN,  // examples/ukf-pen.rs:76-91
    PEN_NL = MPCB_MODEL_PEN_NL,    // examples/ukf-pen2.rs:31-53
    PEN6 = MPCB_MODEL_PEN6,        // examples/ukf-pen3.rs:35-63
    NL6_UKF = MPCB_MODEL_NL6_UKF,  // examples/mppi4-non-liner-ukf.rs:149-179
    USER_UKF = MPCB_MODEL_USER_UKF  // fx / hx supplied as CUDA source (create_user)
};

template <typename T>
struct Result {
    bool ok;
    T value;
    const char* err;  // the reference's &'static str when !ok
    explicit operator bool() const { return ok; }
    const T& unwrap() const {
        if (!ok) throw std::runtime_error(err);
        return value;
    }
};

namespace detail {
inline void check(mpcb_status st, const char* what) {
    if (st != MPCB_OK)
        throw std::runtime_error(std::string(what) + ": " + mpcb_status_string(st) + " (" + mpcb_last_error_string() + ")");
}
}  // namespace detail

// ------------------------------------------------------------------------------------------------ mppi
namespace mppi {

template <std::size_t N, std::size_t K, std::size_t S>
class Mppi {
   public:
    using VecS = std::array<double, S>;
    using VecN = std::array<double, N>;

    // Mppi::new(dynamics, cost, lambda, std_dev, limit) — src/mppi.rs:16-22
    static Mppi create(DeviceModel dynamics, DeviceModel cost, double lambda, double std_dev, std::pair<double, double> limit,
                       double dt = 0.0, int precision = -1, uint64_t seed = 0) {
        if (dynamics != cost) throw std::invalid_argument("dynamics and cost must name the same built-in device model");
        mpcb_mppi_cfg cfg;
        detail::check(mpcb_mppi_default_cfg(static_cast<int32_t>(dynamics), &cfg), "mpcb_mppi_default_cfg");
        cfg.horizon = static_cast<int32_t>(N);
        cfg.samples = static_cast<int64_t>(K);
        cfg.state_dim = static_cast<int32_t>(S);
        cfg.lambda = lambda;
        cfg.std_dev = std_dev;
        cfg.limit_lo = limit.first;
        cfg.limit_hi = limit.second;
        if (dt > 0.0) cfg.model.dt = dt;  // the examples' DT = T / N constant
        if (precision >= 0) cfg.precision = precision;
        if (seed != 0) cfg.seed = seed;
        Mppi m;
        detail::check(mpcb_mppi_create(&m.h_, &cfg), "mpcb_mppi_create");
        return m;
    }

    // Mppi::new with the caller's OWN dynamics and cost (src/mppi.rs:9-10): the two functions as CUDA C++ source
    //     template <typename real> void dynamics(real (&x)[4], real u, const real* p);   // x <- f(x, u)
    //     template <typename real> real cost(const real (&x)[4], const real* p);
    // compiled into the fused kernel at construction (mpcb_mppi_create_user); `params` are the p[] constants.
    static Mppi create_user(const std::string& cuda_source, const std::vector<double>& params, double lambda, double std_dev,
                            std::pair<double, double> limit, int precision = MPCB_F32, uint64_t seed = 0) {
        mpcb_mppi_cfg cfg;
        detail::check(mpcb_mppi_default_cfg(MPCB_MODEL_USER, &cfg), "mpcb_mppi_default_cfg");
        cfg.horizon = static_cast<int32_t>(N);
        cfg.samples = static_cast<int64_t>(K);
        cfg.state_dim = static_cast<int32_t>(S);
        cfg.lambda = lambda;
        cfg.std_dev = std_dev;
        cfg.limit_lo = limit.first;
        cfg.limit_hi = limit.second;
        cfg.precision = precision;
        if (seed != 0) cfg.seed = seed;
        Mppi m;
        const mpcb_status st = mpcb_mppi_create_user(&m.h_, &cfg, cuda_source.c_str(), params.data(), static_cast<int32_t>(params.size()));
        if (st == MPCB_RTC_ERROR) throw std::runtime_error(std::string("user model did not compile: ") + mpcb_last_error_string() + "\n" + mpcb_rtc_log());
        detail::check(st, "mpcb_mppi_create_user");
        return m;
    }

    // compute(&mut self, x, u_n) -> Result<SVector<f64,N>, &'static str> — src/mppi.rs:33-92
    Result<VecN> compute(const VecS& x, const VecN& u_n) {
        Result<VecN> r{false, {}, nullptr};
        return finish(mpcb_mppi_compute(h_, x.data(), u_n.data(), r.value.data(), &info_), r);
    }

    // Verification extension: the same step on caller-supplied noise eps[K][N] ~ N(0, std_dev^2) (the reference
    // draws from OS entropy, src/mppi.rs:41, and cannot be replayed).
    Result<VecN> compute_replay(const VecS& x, const VecN& u_n, const double* eps) {
        Result<VecN> r{false, {}, nullptr};
        return finish(mpcb_mppi_compute_replay(h_, x.data(), u_n.data(), eps, MPCB_DT_F64, 0, r.value.data(), &info_), r);
    }

    const mpcb_mppi_info& info() const { return info_; }  // argmax, max, sum, n_finite of the last compute

    Mppi(Mppi&& o) noexcept : h_(o.h_), info_(o.info_) { o.h_ = nullptr; }
    Mppi& operator=(Mppi&& o) noexcept {
        if (this != &o) {
            mpcb_mppi_destroy(h_);
            h_ = o.h_;
            info_ = o.info_;
            o.h_ = nullptr;
        }
        return *this;
    }
    Mppi(const Mppi&) = delete;  // the reference derives Clone; a handle owns device buffers, so build a second one
    Mppi& operator=(const Mppi&) = delete;
    ~Mppi() { mpcb_mppi_destroy(h_); }

   private:
    Mppi() : h_(nullptr), info_{} {}
    Result<VecN> finish(mpcb_status st, Result<VecN>& r) {
        switch (st) {
            case MPCB_OK: r.ok = true; return r;
            case MPCB_NO_FINITE_COST:  // "Cannot calculate max"  src/mppi.rs:69
            case MPCB_SUM_ZERO:        // "sum is zero"           src/mppi.rs:77
            case MPCB_U_INVALID:       // "u is invalid"          src/mppi.rs:88
                r.err = mpcb_status_string(st);
                return r;
            default: detail::check(st, "mpcb_mppi_compute"); return r;
        }
    }
    mpcb_mppi* h_;
    mpcb_mppi_info info_;
};

}  // namespace mppi

// ------------------------------------------------------------------------------------------------ ukf / ukf2
namespace detail {

template <std::size_t NS, std::size_t NO>
class UkfBase {
   public:
    using State = std::array<double, NS>;
    using Obs = std::array<double, NO>;
    using CovN = std::array<double, NS * NS>;  // row-major
    using CovO = std::array<double, NO * NO>;

    // predict(u, fx) — src/ukf.rs:44-52; dt > 0 replaces the dt the fx closure captures
    // (examples/mppi4-non-liner-ukf.rs:278), else the example's DT constant is used
    void predict(double u, DeviceModel fx, double dt = 0.0) {
        same_model(fx);
        check(mpcb_ukf_predict(h_, nullptr, u, dt), "mpcb_ukf_predict");
        raise_on_failure();
    }
    // update(&z, hx) — src/ukf.rs:54-74; .expect("Inverse fail") becomes std::runtime_error("Inverse fail")
    void update(const Obs& z, DeviceModel hx) {
        same_model(hx);
        check(mpcb_ukf_update(h_, z.data()), "mpcb_ukf_update");
        raise_on_failure();
    }
    State state() const {  // src/ukf.rs:88-90
        State x;
        check(mpcb_ukf_get_state(h_, x.data(), nullptr), "mpcb_ukf_get_state");
        return x;
    }
    CovN covariance() const {  // src/ukf.rs:92-94
        CovN p;
        check(mpcb_ukf_get_state(h_, nullptr, p.data()), "mpcb_ukf_get_state");
        return p;
    }

    UkfBase(UkfBase&& o) noexcept : h_(o.h_), model_(o.model_) { o.h_ = nullptr; }
    UkfBase(const UkfBase&) = delete;
    UkfBase& operator=(const UkfBase&) = delete;
    ~UkfBase() { mpcb_ukf_destroy(h_); }

   protected:
    UkfBase(const State& x, const CovN& p, const CovN& q, const CovO& r, DeviceModel model) : h_(nullptr), model_(model) {
        mpcb_ukf_cfg cfg;
        check(mpcb_ukf_default_cfg(static_cast<int32_t>(model), &cfg), "mpcb_ukf_default_cfg");
        if (cfg.n != static_cast<int32_t>(NS) || cfg.o != static_cast<int32_t>(NO))
            throw std::invalid_argument("device model has other state/observation dimensions than this filter type");
        cfg.batch = 1;
        check(mpcb_ukf_create(&h_, &cfg), "mpcb_ukf_create");
        check(mpcb_ukf_init(h_, x.data(), p.data(), q.data(), r.data()), "mpcb_ukf_init");
    }
    // the caller's own fx / hx closures (src/ukf.rs:44-46,54-56) as CUDA C++ source — mpcb_ukf_create_user:
    //     void fx(double (&x)[N], double u, double dt, const double* p);   void hx(const double (&x)[N], double (&z)[O], const double* p);
    UkfBase(const State& x, const CovN& p, const CovN& q, const CovO& r, const std::string& cuda_source, const std::vector<double>& params)
        : h_(nullptr), model_(DeviceModel::USER_UKF) {
        mpcb_ukf_cfg cfg;
        check(mpcb_ukf_default_cfg(MPCB_MODEL_USER_UKF, &cfg), "mpcb_ukf_default_cfg");
        cfg.n = static_cast<int32_t>(NS);
        cfg.o = static_cast<int32_t>(NO);
        cfg.batch = 1;
        const mpcb_status st = mpcb_ukf_create_user(&h_, &cfg, cuda_source.c_str(), params.data(), static_cast<int32_t>(params.size()));
        if (st == MPCB_RTC_ERROR) throw std::runtime_error(std::string("user model did not compile: ") + mpcb_last_error_string() + "\n" + mpcb_rtc_log());
        check(st, "mpcb_ukf_create_user");
        check(mpcb_ukf_init(h_, x.data(), p.data(), q.data(), r.data()), "mpcb_ukf_init");
    }
    void same_model(DeviceModel m) const {
        if (m != model_) throw std::invalid_argument("fx/hx must be the device model the filter was built with");
    }
    void raise_on_failure() const {
        int32_t s = 0;
        const mpcb_status st = mpcb_ukf_get_status(h_, &s);
        if (st == MPCB_INVERSE_FAIL || st == MPCB_CHOLESKY_FAIL) throw std::runtime_error(mpcb_status_string(st));
        check(st, "mpcb_ukf_get_status");
    }
    mpcb_ukf* h_;
    DeviceModel model_;
};

}  // namespace detail

namespace ukf {
// mpc::ukf::UnscentedKalmanFilter — n = 4, o = 3 (src/ukf.rs:1-133)
class UnscentedKalmanFilter : public detail::UkfBase<4, 3> {
   public:
    // UnscentedKalmanFilter::new(x, p, q, r) — src/ukf.rs:30-42 (+ the device model standing in for fx/hx)
    static UnscentedKalmanFilter create(const State& x, const CovN& p, const CovN& q, const CovO& r,
                                        DeviceModel model = DeviceModel::PEN_NL) {
        return UnscentedKalmanFilter(x, p, q, r, model);
    }
    // new(x, p, q, r) for a filter whose predict/update get the caller's own fx / hx (CUDA source, params = p[])
    static UnscentedKalmanFilter create_user(const State& x, const CovN& p, const CovN& q, const CovO& r, const std::string& cuda_source,
                                             const std::vector<double>& params = {}) {
        return UnscentedKalmanFilter(x, p, q, r, cuda_source, params);
    }

   private:
    using UkfBase::UkfBase;
};
}  // namespace ukf

namespace ukfn {
// The same filter for any dimensions (n = NS in 1..6, o = NO in 1..5) with the caller's own fx / hx closures given as CUDA
// C++ source (mpcb_ukf_create_user) — the reference hard-codes (4,3) in mpc::ukf and (6,5) in mpc::ukf2; its algorithm,
// which is what the kernel implements, does not depend on them.
template <std::size_t NS, std::size_t NO>
class UnscentedKalmanFilter : public detail::UkfBase<NS, NO> {
    using Base = detail::UkfBase<NS, NO>;

   public:
    static UnscentedKalmanFilter create_user(const typename Base::State& x, const typename Base::CovN& p, const typename Base::CovN& q,
                                             const typename Base::CovO& r, const std::string& cuda_source,
                                             const std::vector<double>& params = {}) {
        return UnscentedKalmanFilter(x, p, q, r, cuda_source, params);
    }
    void set_q(const typename Base::CovN& q) { detail::check(mpcb_ukf_set_q(this->h_, q.data()), "mpcb_ukf_set_q"); }
    void set_r(const typename Base::CovO& r) { detail::check(mpcb_ukf_set_r(this->h_, r.data()), "mpcb_ukf_set_r"); }

   private:
    using Base::Base;
};
}  // namespace ukfn

namespace ukf2 {
// mpc::ukf2::UnscentedKalmanFilter — n = 6, o = 5 (src/ukf2.rs:1-137)
class UnscentedKalmanFilter : public detail::UkfBase<6, 5> {
   public:
    static UnscentedKalmanFilter create(const State& x, const CovN& p, const CovN& q, const CovO& r,
                                        DeviceModel model = DeviceModel::NL6_UKF) {
        return UnscentedKalmanFilter(x, p, q, r, model);
    }
    static UnscentedKalmanFilter create_user(const State& x, const CovN& p, const CovN& q, const CovO& r, const std::string& cuda_source,
                                             const std::vector<double>& params = {}) {
        return UnscentedKalmanFilter(x, p, q, r, cuda_source, params);
    }
    void set_q(const CovN& q) { detail::check(mpcb_ukf_set_q(h_, q.data()), "mpcb_ukf_set_q"); }  // src/ukf2.rs:96-98
    // called by examples/mppi4-ukf-commu.rs:280 but missing in the reference
    void set_r(const CovO& r) { detail::check(mpcb_ukf_set_r(h_, r.data()), "mpcb_ukf_set_r"); }
    // the hx closure of examples/mppi4-ukf-commu.rs:279-293 zeroes the rows of disabled sensors: hand over the mask
    void set_enable(uint8_t enable) { detail::check(mpcb_ukf_set_enable(h_, enable), "mpcb_ukf_set_enable"); }
    // gen_r(enable) — examples/mppi4-ukf-commu.rs:228-236
    CovO gen_r(uint8_t enable, const CovO& r) const {
        CovO out;
        detail::check(mpcb_ukf_gen_r(h_, enable, r.data(), out.data()), "mpcb_ukf_gen_r");
        return out;
    }

   private:
    using UkfBase::UkfBase;
};
}  // namespace ukf2

// ------------------------------------------------------------------------------------------------ gaussian
namespace gaussian {
// src/gaussian.rs:1-63 — host-side scalar type, no GPU content
struct Gaussian {
    double mean = 0.0;  // Default: (0, 0) — :13-20
    double var = 0.0;
    static Gaussian create(double mean, double var) { return Gaussian{mean, var}; }  // :8-10
};
inline Gaussian operator+(Gaussian a, Gaussian b) { return {a.mean + b.mean, a.var + b.var}; }  // :22-30
inline Gaussian operator-(Gaussian a, Gaussian b) { return {a.mean - b.mean, a.var - b.var}; }  // :32-40 (var subtracts)
inline Gaussian operator*(Gaussian a, Gaussian b) {                                             // :42-52
    return {(a.var * b.mean + b.var * a.mean) / (a.var + b.var), (a.var * b.var) / (a.var + b.var)};
}
inline Gaussian operator*(Gaussian a, double s) { return {a.mean * s, a.var * s}; }  // :54-62 (scales mean AND var)
}  // namespace gaussian

}  // namespace mpc

#endif  // MPC_B200_HPP
