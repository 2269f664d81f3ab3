// examples/ukf-pen2.rs on B200 through the C++ mirror: the library UKF (mpc::ukf, n = 4, o = 3) tracking the nonlinear
// pendulum for 100 steps, same constants and printed line (examples/ukf-pen2.rs:8-110).  fx/hx are the PEN_NL device
// model; the truth model and the simulated sensor stay on the host like in the reference.
//   g++ -std=c++17 -O2 -Iinclude examples/cpp/ukf_pen2.cpp -Lmpc_rs_b200 -lmpc_b200 -Wl,-rpath,$PWD/mpc_rs_b200 -o examples/cpp/ukf_pen2
#include <cmath>
#include <cstdio>
#include <random>

#include "mpc_b200.hpp"

using Vec4 = std::array<double, 4>;
constexpr double M1 = 150e-3, R_W = 50e-3, M2 = 2.3 - 2.0 * M1 + 2.0, L = 0.2474, J1 = M1 * R_W * R_W, J2 = 0.2, G = 9.81, KT = 0.15;
constexpr double DT = 0.01;

// examples/ukf-pen2.rs:31-44
static Vec4 fx(const Vec4& x, double u) {
    Vec4 r = x;
    constexpr double D = (M1 + M2 + J1 / (R_W * R_W)) * (M2 * L * L + J2);
    const double d = D - M2 * M2 * L * L * std::cos(x[2]) * std::cos(x[2]);
    const double term1 = (M1 + M2 + J1 / (R_W * R_W)) * M2 * G * L * std::sin(x[2]);
    const double term2 = (KT * u / R_W + M2 * L * x[3] * x[3] * std::sin(x[2])) * M2 * L * std::cos(x[2]);
    r[3] += (term1 - term2) / d * DT;
    r[2] += x[3] * DT;
    const double term3 = (J2 + M2 * L * L) * (KT * u / R_W + M2 * L * x[3] * x[3] * std::sin(x[2]));
    const double term4 = M2 * G * L * L * std::sin(x[2]) * std::cos(x[2]);
    r[1] += (term3 + term4) / d * DT;
    r[0] += x[1] * DT;
    return r;
}
// examples/ukf-pen2.rs:47-53
static std::array<double, 3> hx(const Vec4& x) {
    return {60.0 / (2.0 * M_PI * R_W) * x[1], 60.0 / (2.0 * M_PI * R_W) * x[1], x[3] * 180.0 / M_PI};
}

int main(int argc, char** argv) {
    const int steps = argc > 1 ? std::atoi(argv[1]) : 100;
    std::mt19937_64 rng(argc > 2 ? std::atoll(argv[2]) : 1);
    std::normal_distribution<double> dist(0.0, 1.0);
    const std::array<double, 16> Q{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.25};  // :18-23
    const std::array<double, 9> R{100, 0, 0, 0, 100, 0, 0, 0, 0.5};                       // :24-28
    const std::array<double, 16> P{10, 0, 0, 0, 0, 10, 0, 0, 0, 0, 10, 0, 0, 0, 0, 10};  // :71-76
    Vec4 x_act{0, 0, 0, 0};
    auto ukf = mpc::ukf::UnscentedKalmanFilter::create({0, 0, 0, 0}, P, Q, R, mpc::DeviceModel::PEN_NL);
    for (int i = 0; i < steps; ++i) {
        const double u = 0.1;
        x_act = fx(x_act, u);
        ukf.predict(u, mpc::DeviceModel::PEN_NL);
        auto x_obs = hx(x_act);  // sensor(): hx + (100, 100, 0.5) * N(0,1), :56-66
        x_obs[0] += 100.0 * dist(rng);
        x_obs[1] += 100.0 * dist(rng);
        x_obs[2] += 0.5 * dist(rng);
        ukf.update(x_obs, mpc::DeviceModel::PEN_NL);
        const auto x_est = ukf.state();
        const auto p = ukf.covariance();
        std::printf("t: %4.2f x_act: (%7.2f,%7.2f,%7.2f,%7.2f) x_obs: (%7.2f,%7.2f) x_est: (%7.2f,%7.2f,%7.2f,%7.2f) p: (%7.2f,%7.2f,%7.2f,%7.2f)\n",
                    i * DT, x_act[0], x_act[1], x_act[2], x_act[3], x_obs[0], x_obs[1], x_est[0], x_est[1], x_est[2], x_est[3], p[0],
                    p[5], p[10], p[15]);
    }
    return 0;
}
