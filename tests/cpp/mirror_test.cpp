// Parity of the C++ host mirror (include/mpc_b200.hpp) against the CPU oracle (oracle/, test infrastructure): written
// the way a test of the reference crate would read — construct with the reference's constants, call compute /
// predict / update, compare.  Built and run by tests/test_cpp_mirror.py (-m gpu); exits 0 and prints "mirror ok".
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "mpc_b200.hpp"
#include "mpc_oracle.h"

static int fails = 0;
#define EXPECT(cond, ...)                          \
    do {                                           \
        if (!(cond)) {                             \
            std::printf("FAIL %s:%d: ", __FILE__, __LINE__); \
            std::printf(__VA_ARGS__);              \
            std::printf("\n");                     \
            ++fails;                               \
        }                                          \
    } while (0)

template <std::size_t N>
static double rel_err(const std::array<double, N>& a, const double* b) {
    double num = 0, den = 0;
    for (std::size_t i = 0; i < N; ++i) { num += (a[i] - b[i]) * (a[i] - b[i]); den += b[i] * b[i]; }
    return std::sqrt(num) / std::sqrt(den > 0 ? den : 1e-300);
}

int main() {
    using namespace mpc;
    // ---- Mppi<8, 20000, 4>, model L, shipped constants (examples/mppi4.rs:8-18), 3 closed-loop steps on replayed noise
    {
        constexpr std::size_t N = 8, K = 20000;
        auto m64 = mppi::Mppi<N, K, 4>::create(DeviceModel::L, DeviceModel::L, 0.5, 3.0, {-20.0, 20.0}, 0.0, MPCB_F64);
        auto m32 = mppi::Mppi<N, K, 4>::create(DeviceModel::L, DeviceModel::L, 0.5, 3.0, {-20.0, 20.0});
        mpcb_model_params p;
        orc_model_defaults(MPCB_MODEL_L, &p);
        std::mt19937_64 rng(7);
        std::normal_distribution<double> nd(0.0, 3.0);
        std::array<double, 4> x{0.5, 0.0, 0.1, 0.0};
        std::array<double, N> u{};
        std::vector<double> eps(K * N);
        for (int step = 0; step < 3; ++step) {
            for (auto& e : eps) e = nd(rng);
            double u_ref[N];
            orc_mppi_out info;
            const int st = orc_mppi_compute(MPCB_MODEL_L, &p, K, N, 0.5, 3.0, -20.0, 20.0, x.data(), u.data(), eps.data(), u_ref, nullptr, &info);
            EXPECT(st == 0, "oracle status %d", st);
            const auto r64 = m64.compute_replay(x, u, eps.data());
            const auto r32 = m32.compute_replay(x, u, eps.data());
            EXPECT(r64.ok && r32.ok, "compute failed: %s", r64.ok ? r32.err : r64.err);
            EXPECT(rel_err(r64.value, u_ref) < 1e-9, "f64 controls %.3e", rel_err(r64.value, u_ref));
            EXPECT(rel_err(r32.value, u_ref) < 1e-5, "f32 controls %.3e", rel_err(r32.value, u_ref));
            EXPECT(m64.info().argmax == info.argmax && m32.info().argmax == info.argmax, "argmin");
            double xn[4];
            orc_dynamics(MPCB_MODEL_L, &p, x.data(), u_ref[0], xn);
            std::memcpy(x.data(), xn, sizeof(xn));
            std::memcpy(u.data(), u_ref, sizeof(u_ref));
        }
        // generate mode returns a usable control (what examples/mppi4.rs:42 does with .unwrap())
        const auto g = m32.compute(x, u).unwrap();
        EXPECT(std::isfinite(g[0]) && std::fabs(g[0]) <= 20.0, "generate-mode control %f", g[0]);
        // Err(&'static str): NaN state -> no finite cost -> "Cannot calculate max" (src/mppi.rs:69)
        const auto bad = m64.compute({NAN, 0, 0, 0}, u);
        EXPECT(!bad.ok && std::strcmp(bad.err, "Cannot calculate max") == 0, "error path: %s", bad.err ? bad.err : "(ok)");
        bool threw = false;
        try { bad.unwrap(); } catch (const std::runtime_error& e) { threw = std::strcmp(e.what(), "Cannot calculate max") == 0; }
        EXPECT(threw, "unwrap() must throw the reference's message");
    }
    // ---- Mppi::create_user: examples/mppi4.rs's dynamics and cost handed over as source (the fn pointers of Mppi::new,
    //      src/mppi.rs:9-10) must reproduce the oracle's model L
    {
        constexpr std::size_t N = 8, K = 8192;
        const std::string src = R"SRC(
template <typename real> void dynamics(real (&x)[4], real u, const real* p) {
    // p = a1, b1, a2, b2, DT  (examples/mppi4.rs:81-88: semi-implicit Euler, x3 x2 x1 x0 in that order)
    x[3] += (p[0] * x[2] - p[1] * u) * p[4];
    x[2] += x[3] * p[4];
    x[1] += (p[2] * x[2] + p[3] * u) * p[4];
    x[0] += x[1] * p[4];
}
template <typename real> real cost(const real (&x)[4], const real* p) {
    const real xc = mpcb::clampm(x[0], (real)-2.0, (real)2.0);
    const real a = mpcb::clampm(x[1] + (real)2.0 * xc, (real)-5.0, (real)5.0);
    const real b = x[2] + (real)0.35 * mpcb::clampm(x[0], (real)-0.75, (real)0.75);
    return (real)2.0 * (xc * xc) + (real)3.0 * (a * a) + (real)5.0 * (b * b) + (real)1.2 * (x[3] * x[3]);
}
)SRC";
        mpcb_model_params p;
        orc_model_defaults(MPCB_MODEL_L, &p);
        const double D = (p.m1 + p.m2 + p.j1 / (p.r_w * p.r_w)) * (p.m2 * p.l * p.l + p.j2) - p.m2 * p.m2 * p.l * p.l;
        const std::vector<double> prm{(p.m1 + p.m2 + p.j1 / (p.r_w * p.r_w)) / D * p.m2 * p.g * p.l, p.m2 * p.l / D / p.r_w * p.kt,
                                      -p.m2 * p.m2 * p.g * p.l * p.l / D, (p.m2 * p.l * p.l + p.j2) / D / p.r_w * p.kt, p.dt};
        auto mu = mppi::Mppi<N, K, 4>::create_user(src, prm, 0.5, 3.0, {-20.0, 20.0}, MPCB_F64);
        std::mt19937_64 rng(17);
        std::normal_distribution<double> nd(0.0, 3.0);
        const std::array<double, 4> x{0.5, 0.0, 0.1, 0.0};
        std::array<double, N> u{};
        std::vector<double> eps(K * N);
        for (auto& e : eps) e = nd(rng);
        double u_ref[N];
        orc_mppi_out info;
        EXPECT(orc_mppi_compute(MPCB_MODEL_L, &p, K, N, 0.5, 3.0, -20.0, 20.0, x.data(), u.data(), eps.data(), u_ref, nullptr, &info) == 0, "oracle");
        const auto r = mu.compute_replay(x, u, eps.data());
        EXPECT(r.ok && rel_err(r.value, u_ref) < 1e-9 && mu.info().argmax == info.argmax, "user model L: %.3e", r.ok ? rel_err(r.value, u_ref) : -1.0);
        bool threw = false;
        try { (void)mppi::Mppi<N, K, 4>::create_user("void dynamics() {}", {}, 0.5, 3.0, {-1.0, 1.0}); }
        catch (const std::runtime_error& e) { threw = std::strstr(e.what(), "did not compile") != nullptr; }
        EXPECT(threw, "a user model that does not compile must throw with the compiler log");
    }
    // ---- mpc::ukf::UnscentedKalmanFilter (n=4, o=3), examples/ukf-pen2.rs constants, 5 predict/update pairs vs the oracle
    {
        const std::array<double, 16> Q{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.25};
        const std::array<double, 9> R{100, 0, 0, 0, 100, 0, 0, 0, 0.5};
        const std::array<double, 16> P0{10, 0, 0, 0, 0, 10, 0, 0, 0, 0, 10, 0, 0, 0, 0, 10};
        auto f = ukf::UnscentedKalmanFilter::create({0, 0, 0, 0}, P0, Q, R, DeviceModel::PEN_NL);
        mpcb_model_params p;
        orc_model_defaults(MPCB_MODEL_PEN_NL, &p);
        double xo[4] = {0, 0, 0, 0}, Po[16], sf[4 * 9];
        std::memcpy(Po, P0.data(), sizeof(Po));
        std::mt19937_64 rng(3);
        std::normal_distribution<double> nd(0.0, 1.0);
        for (int i = 0; i < 5; ++i) {
            const std::array<double, 3> z{30 * nd(rng), 30 * nd(rng), 0.5 * nd(rng)};
            EXPECT(orc_ukf_predict(MPCB_MODEL_PEN_NL, &p, 4, MPCB_SQRT_EIG, MPCB_ORDER_LIBRARY, xo, Po, Q.data(), 0.1, 0.0, sf) == 0, "oracle predict");
            EXPECT(orc_ukf_update(MPCB_MODEL_PEN_NL, &p, 4, 3, xo, Po, R.data(), z.data(), sf) == 0, "oracle update");
            f.predict(0.1, DeviceModel::PEN_NL);
            f.update(z, DeviceModel::PEN_NL);
            // the filters are re-synchronised every step: the sigma weights amplify rounding 1.7e5 x per step
            EXPECT(rel_err(f.state(), xo) < 1e-6, "ukf state step %d: %.3e", i, rel_err(f.state(), xo));
            EXPECT(rel_err(f.covariance(), Po) < 1e-6, "ukf covariance step %d: %.3e", i, rel_err(f.covariance(), Po));
        }
        bool threw = false;
        try { f.predict(0.1, DeviceModel::PEN6); } catch (const std::invalid_argument&) { threw = true; }
        EXPECT(threw, "a foreign fx model must be rejected");
    }
    // ---- ukf::UnscentedKalmanFilter::create_user: examples/ukf-pen2.rs's fx / hx closures as source vs the oracle's PEN_NL
    {
        const std::string src = R"SRC(
void fx(double (&x)[4], double u, double dt, const double* p) {
    const double M1 = p[0], R_W = p[1], M2 = p[2], L = p[3], J1 = p[4], J2 = p[5], G = p[6], KT = p[7];
    const double s = sin(x[2]), c = cos(x[2]);
    const double D = (M1 + M2 + J1 / (R_W * R_W)) * (M2 * L * L + J2);
    const double d = D - M2 * M2 * L * L * c * c;
    const double drive = KT * u / R_W + M2 * L * (x[3] * x[3]) * s;
    const double r3 = x[3] + ((M1 + M2 + J1 / (R_W * R_W)) * M2 * G * L * s - drive * M2 * L * c) / d * dt;
    const double r2 = x[2] + x[3] * dt;
    const double r1 = x[1] + ((J2 + M2 * L * L) * drive + M2 * G * L * L * s * c) / d * dt;
    const double r0 = x[0] + x[1] * dt;
    x[0] = r0; x[1] = r1; x[2] = r2; x[3] = r3;
}
void hx(const double (&x)[4], double (&z)[3], const double* p) {
    const double PI = 3.14159265358979323846264338327950288;
    z[0] = 60.0 / (2.0 * PI * p[1]) * x[1];
    z[1] = 60.0 / (2.0 * PI * p[1]) * x[1];
    z[2] = x[3] * (180.0 / PI);
}
)SRC";
        mpcb_model_params p;
        orc_model_defaults(MPCB_MODEL_PEN_NL, &p);
        const std::array<double, 16> Q{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.25};
        const std::array<double, 9> R{100, 0, 0, 0, 100, 0, 0, 0, 0.5};
        const std::array<double, 16> P0{10, 0, 0, 0, 0, 10, 0, 0, 0, 0, 10, 0, 0, 0, 0, 10};
        auto f = ukf::UnscentedKalmanFilter::create_user({0.01, 0, 0.02, 0}, P0, Q, R, src, {p.m1, p.r_w, p.m2, p.l, p.j1, p.j2, p.g, p.kt});
        double xo[4] = {0.01, 0, 0.02, 0}, Po[16], sf[4 * 9];
        std::memcpy(Po, P0.data(), sizeof(Po));
        const std::array<double, 3> z{12.0, -7.0, 0.3};
        EXPECT(orc_ukf_predict(MPCB_MODEL_PEN_NL, &p, 4, MPCB_SQRT_EIG, MPCB_ORDER_LIBRARY, xo, Po, Q.data(), 0.1, 0.01, sf) == 0, "oracle predict");
        EXPECT(orc_ukf_update(MPCB_MODEL_PEN_NL, &p, 4, 3, xo, Po, R.data(), z.data(), sf) == 0, "oracle update");
        f.predict(0.1, DeviceModel::USER_UKF, 0.01);
        f.update(z, DeviceModel::USER_UKF);
        EXPECT(rel_err(f.state(), xo) < 1e-6 && rel_err(f.covariance(), Po) < 1e-6, "user ukf: %.3e %.3e", rel_err(f.state(), xo),
               rel_err(f.covariance(), Po));
    }
    // ---- ukfn::UnscentedKalmanFilter<3, 2>: a filter of other dimensions with the caller's own fx / hx; a linear model,
    //      so that the unscented filter has a closed form: the Kalman filter of examples/two-liner-kf.rs:17-52, except
    //      that the reference's update reuses the sigma points propagated by predict (src/ukf.rs:54-74), which do not
    //      carry Q — innovation covariance and cross covariance are built from F P F^T, the covariance update from
    //      F P F^T + Q
    {
        const std::string src = R"SRC(
void fx(double (&x)[3], double u, double dt, const double* p) { x[0] += x[1] * dt; x[1] += (u + x[2]) * dt; }
void hx(const double (&x)[3], double (&z)[2], const double* p) { z[0] = x[0]; z[1] = x[1] + p[0] * x[2]; }
)SRC";
        const double dt = 0.1, u = 0.7, c = 0.5;
        const std::array<double, 9> Q{1e-3, 0, 0, 0, 1e-2, 0, 0, 0, 1e-4}, P0{1, 0.1, 0, 0.1, 2, 0.2, 0, 0.2, 3};
        const std::array<double, 4> R{0.05, 0, 0, 0.2};
        auto f = ukfn::UnscentedKalmanFilter<3, 2>::create_user({0.3, -0.1, 0.05}, P0, Q, R, src, {c});
        // closed-form KF with F = [[1,dt,0],[0,1,dt],[0,0,1]], B = [0,dt,0], H = [[1,0,0],[0,1,c]]
        double x[3] = {0.3, -0.1, 0.05}, P[3][3] = {{1, 0.1, 0}, {0.1, 2, 0.2}, {0, 0.2, 3}};
        const double F[3][3] = {{1, dt, 0}, {0, 1, dt}, {0, 0, 1}}, Hm[2][3] = {{1, 0, 0}, {0, 1, c}};
        const std::array<double, 2> z{0.25, 0.1};
        {
            double xp[3] = {x[0] + x[1] * dt, x[1] + (u + x[2]) * dt, x[2]}, FP[3][3] = {}, Pp[3][3] = {};
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) FP[i][j] += F[i][k] * P[k][j];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { for (int k = 0; k < 3; ++k) Pp[i][j] += FP[i][k] * F[j][k]; Pp[i][j] += Q[i * 3 + j]; }
            double S[2][2] = {}, PHt[3][2] = {}, Pm[3][3];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Pm[i][j] = Pp[i][j] - Q[i * 3 + j];  // F P F^T
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 2; ++j) for (int k = 0; k < 3; ++k) PHt[i][j] += Pm[i][k] * Hm[j][k];
            for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) { for (int k = 0; k < 3; ++k) S[i][j] += Hm[i][k] * PHt[k][j]; S[i][j] += R[i * 2 + j]; }
            const double det = S[0][0] * S[1][1] - S[0][1] * S[1][0];
            const double Si[2][2] = {{S[1][1] / det, -S[0][1] / det}, {-S[1][0] / det, S[0][0] / det}};
            double K[3][2] = {};
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 2; ++j) for (int k = 0; k < 2; ++k) K[i][j] += PHt[i][k] * Si[k][j];
            const double inn[2] = {z[0] - xp[0], z[1] - (xp[1] + c * xp[2])};
            for (int i = 0; i < 3; ++i) x[i] = xp[i] + K[i][0] * inn[0] + K[i][1] * inn[1];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double kskt = 0; for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) kskt += K[i][a] * S[a][b] * K[j][b]; P[i][j] = Pp[i][j] - kskt; }
        }
        f.predict(u, DeviceModel::USER_UKF, dt);
        f.update(z, DeviceModel::USER_UKF);
        const auto xs = f.state();
        const auto Ps = f.covariance();
        double ex = 0, ep = 0;
        for (int i = 0; i < 3; ++i) ex = std::fmax(ex, std::fabs(xs[i] - x[i]));
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) ep = std::fmax(ep, std::fabs(Ps[i * 3 + j] - P[i][j]));
        // exact for linear models up to the sigma-weight amplification of rounding (1.7e5 x 1e-16 per operation)
        EXPECT(ex < 1e-8 && ep < 1e-8, "linear-model UKF<3,2> vs closed-form KF: %.3e %.3e", ex, ep);
    }
    // ---- mpc::ukf2 (n=6, o=5): set_q / set_r / set_enable / gen_r exist and a gated step runs
    {
        std::array<double, 36> Q{}, P0{};
        std::array<double, 25> R{};
        mpcb_ukf_default_noise(MPCB_MODEL_NL6_UKF, 0.01, Q.data(), R.data(), P0.data());
        auto f = ukf2::UnscentedKalmanFilter::create({0, 0, 0, 0.05, 0, 0}, P0, Q, R);
        f.set_q(Q);
        const auto Rg = f.gen_r(0b10101, R);
        EXPECT(Rg[6] == 1e6 && Rg[18] == 1e6 && Rg[0] == R[0], "gen_r");
        f.set_r(Rg);
        f.set_enable(0b10101);
        f.predict(0.3, DeviceModel::NL6_UKF, 0.01);
        f.update({10.0, 999.0, 1.0, 999.0, 0.0}, DeviceModel::NL6_UKF);
        const auto x = f.state();
        EXPECT(std::isfinite(x[0]) && std::isfinite(x[3]), "gated ukf2 step");
    }
    // ---- Gaussian (src/gaussian.rs)
    {
        using gaussian::Gaussian;
        const Gaussian a = Gaussian::create(1.0, 4.0), b = Gaussian::create(3.0, 1.0);
        const Gaussian s = a + b, d = a - b, m = a * b, k = a * 2.0, z{};
        EXPECT(s.mean == 4.0 && s.var == 5.0 && d.mean == -2.0 && d.var == 3.0, "add/sub");
        EXPECT(std::fabs(m.mean - (4.0 * 3.0 + 1.0 * 1.0) / 5.0) < 1e-15 && std::fabs(m.var - 0.8) < 1e-15, "product");
        EXPECT(k.mean == 2.0 && k.var == 8.0 && z.mean == 0.0 && z.var == 0.0, "scale/default");
    }
    if (fails == 0) std::printf("mirror ok\n");
    return fails == 0 ? 0 : 1;
}
