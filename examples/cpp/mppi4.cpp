// examples/mppi4.rs on B200 through the C++ mirror (include/mpc_b200.hpp): same constants, same control loop, same
// printed line and CSV record (examples/mppi4.rs:8-70).  Only the constructor differs: DeviceModel tags instead of fn
// pointers.  The plant step `dynamics(&x, u_n[0])` stays on the host, exactly as in the reference.
//   g++ -std=c++17 -O2 -Iinclude examples/cpp/mppi4.cpp -Lmpc_rs_b200 -lmpc_b200 -Wl,-rpath,$PWD/mpc_rs_b200 -o examples/cpp/mppi4
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>

#include "mpc_b200.hpp"

// examples/mppi4.rs:8-18
constexpr double T = 0.8;
constexpr std::size_t N = 8;
constexpr double DT = T / N;
constexpr std::size_t K = 800000;
constexpr double LAMBDA = 0.5;
constexpr double R = 3.0;
constexpr std::pair<double, double> LIMIT{-20.0, 20.0};

using Vec4 = std::array<double, 4>;

// examples/mppi4.rs:73-89
static Vec4 dynamics(const Vec4& x, double u) {
    constexpr double M1 = 150e-3, R_W = 50e-3, M2 = 2.3 - 2.0 * M1 + 2.0, L = 0.2474, J1 = M1 * R_W * R_W, J2 = 0.2, G = 9.81,
                     KT = 0.15;
    constexpr double D = (M1 + M2 + J1 / (R_W * R_W)) * (M2 * L * L + J2) - M2 * M2 * L * L;
    Vec4 r = x;
    r[3] += ((M1 + M2 + J1 / (R_W * R_W)) / D * M2 * G * L * r[2] - M2 * L / D / R_W * KT * u) * DT;
    r[2] += r[3] * DT;
    r[1] += (-M2 * M2 * G * L * L / D * r[2] + (M2 * L * L + J2) / D / R_W * KT * u) * DT;
    r[0] += r[1] * DT;
    return r;
}

// Rust's f64::to_string: shortest digits that round-trip, no exponent
static std::string rust_f64(double v) {
    char buf[64];
    for (int prec = 1; prec <= 17; ++prec) {
        std::snprintf(buf, sizeof(buf), "%.*g", prec, v);
        if (std::strtod(buf, nullptr) == v) break;
    }
    std::string s(buf);
    if (s.find('e') != std::string::npos) {  // expand the exponent form
        std::snprintf(buf, sizeof(buf), "%.20f", v);
        s = buf;
        while (!s.empty() && s.back() == '0') s.pop_back();
        if (!s.empty() && s.back() == '.') s.pop_back();
    }
    return s;
}

int main(int argc, char** argv) {
    const double seconds = argc > 1 ? std::atof(argv[1]) : 10.0;
    const char* file_path = argc > 2 ? argv[2] : "logs/mppi/mppi.csv";
    Vec4 x{0.5, 0.0, 0.1, 0.0};
    std::array<double, N> u_n{};

    auto mppi = mpc::mppi::Mppi<N, K, 4>::create(mpc::DeviceModel::L, mpc::DeviceModel::L, LAMBDA, R, LIMIT);

    std::system("mkdir -p logs/mppi");
    std::ofstream wtr(file_path);
    const auto now = std::chrono::steady_clock::now();
    double t = 0.0;
    while (t < seconds) {
        u_n = mppi.compute(x, u_n).unwrap();
        x = dynamics(x, u_n[0]);
        std::printf("t: %.2f, u: %6.2f, x: [%6.2f, %5.2f, %5.2f, %5.2f]\n", t, u_n[0], x[0], x[1], x[2], x[3]);
        if (std::fabs(x[2]) > 60.0 * M_PI / 180.0) {
            std::printf("x[2] is over 60 degrees\n");
            break;
        }
        wtr << rust_f64(t) << ',' << rust_f64(u_n[0]) << ',' << rust_f64(x[0]) << ',' << rust_f64(x[1]) << ',' << rust_f64(x[2]) << ','
            << rust_f64(x[3]) << '\n';
        wtr.flush();
        t += DT;
    }
    std::printf("elapsed: %.2f sec\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - now).count());
    return 0;
}
