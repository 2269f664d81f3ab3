// closed_loop.cu — BASELINE config #4 behind the C ABI: C independent robots of examples/mppi4-non-liner-ukf.rs, each
// with its own MPPI controller (model NL6) and its own UKF (model NL6_UKF), one fixed-length tick per call, everything
// on the device:
//
//   plant   x <- dynamics_short(x, u_n[0], tick_dt, push(t))     (:149-159, :236-244)    closed_loop_plant_kernel
//   sensor  z <- hx(x) + R * N(0, 1)                             (:169-190)              the same kernel (Philox noise)
//   UKF     set_q(gen_q(dt)); predict(u_n[0], fx); update(z, hx) (:272-283)              mpcb_ukf_run_device (fused)
//   MPPI    x_est -> [x0, x1, x3, x4]; u_n <- compute(x_est, u_n) (:55-87)               mpcb_mppi_compute_device
//
// The reference runs these as four free-running threads against the wall clock; a batch needs a deterministic schedule
// (fixed tick of 0.01 s).  The plant stands in for the real robot; in round 1 it ran on the host in numpy
// (35 % of a tick).  The two streams (UKF + plant, MPPI) are chained with events, so n ticks are n x 5 asynchronous
// launches and no host round trip.  Controllers shard over GPUs without any exchange: rank r owns the robots
// [controller_offset, controller_offset + C) and its noise counters use the GLOBAL robot index, so the trajectories do
// not depend on the sharding.
#include <math_constants.h>

#include <new>

#include "common.cuh"
#include "philox.cuh"

namespace mpcb {

struct PlantParams {
    double* x;         // [6][C] truth, SoA
    const double* u0;  // [C] control being applied
    double* z;         // [5][C] sensor readings out (SoA, what mpcb_ukf_run_device takes)
    long long C, c_offset;
    double dt, f;      // tick length, push force of this tick
    unsigned int seed_lo, seed_hi, tick;
    int write_z;       // 0: z is supplied by the caller (replay)
    // physical constants (examples/mppi4-non-liner-ukf.rs:108-124) and the sensor's standard deviations (:28)
    double M1, R_W, M2, L, J1, J2, G, KT;
    double r_std[5];
};

// ddot of examples/mppi4-non-liner-ukf.rs:126-139 in the reference's association order (this TU is compiled with
// -fmad=false), including the x[3].cos() quirk of the f terms
__device__ __forceinline__ void plant_ddot(const PlantParams& p, double th, double thd, double u, double f, double* ddx, double* ddth) {
    const double B = p.M2 * p.L * p.L + p.J2;
    const double A2 = 2.0 * p.M1 + p.M2 + 2.0 * p.J1 / (p.R_W * p.R_W);
    const double D1 = A2 * B;
    const double s = sin(th), c = cos(th);
    const double mlc = p.M2 * p.L * c;
    const double d = D1 - mlc * mlc;
    const double w2 = thd * thd;
    const double cf = cos(thd);
    const double term1 = B * p.M2 * p.L / d * w2 * s;
    const double term2 = -(p.M2 * p.L) * (p.M2 * p.L) * p.G / d * s * c;
    const double term3 = 2.0 * B / (d * p.R_W) * p.KT * u;
    const double term4 = B / d * f * cf;
    *ddx = term1 + term2 + term3 + term4;
    const double t1 = -(p.M2 * p.L) * (p.M2 * p.L) / d * w2 * s * c;
    const double t2 = (p.M2 * p.G * s - 2.0 * f) * p.L * A2 / d;
    const double t3 = -2.0 * p.M2 * p.L / (d * p.R_W) * p.KT * u * c;
    const double t4 = -p.M2 * p.L * f * (cf * cf) / d;
    *ddth = t1 + t2 + t3 + t4;
}

__global__ void closed_loop_plant_kernel(const PlantParams p) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.C) return;
    double x[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = p.x[(long long)i * p.C + c];
    // dynamics_short (:149-159): accelerations from ddot on [x0, x1, x3, x4], then semi-implicit Euler
    double ddx, ddth;
    plant_ddot(p, x[3], x[4], p.u0[c], p.f, &ddx, &ddth);
    x[5] = ddth;
    x[4] += x[5] * p.dt;
    x[3] += x[4] * p.dt;
    x[2] = ddx;
    x[1] += x[2] * p.dt;
    x[0] += x[1] * p.dt;
#pragma unroll
    for (int i = 0; i < 6; ++i) p.x[(long long)i * p.C + c] = x[i];
    if (!p.write_z) return;
    // hx (:169-179) + R * N(0,1) (:180-190); noise: Philox counter = (global robot, tick, tag, j)
    const double kPi = 3.14159265358979323846;
    const double ax = p.G * sin(x[3]) + x[2] * cos(x[3]) + p.L * x[5];
    const double az = p.G * cos(x[3]) - x[2] * sin(x[3]) + p.L * (x[4] * x[4]);
    double z[5];
    z[0] = 36.0 * 60.0 / (2.0 * kPi * p.R_W) * x[1];
    z[1] = 36.0 * -60.0 / (2.0 * kPi * p.R_W) * x[1];
    z[2] = x[4] * (180.0 / kPi);
    z[3] = az / p.G;
    z[4] = ax / p.G;
    const unsigned long long g = (unsigned long long)(p.c_offset + c);
    float n[8];
    {
        float a[4], b[4];
        const float one = (float)(-2.0 * 0.693147180559945309417);  // sigma = 1
        philox_normal4(philox4x32((unsigned int)g, p.tick, 0x00C105EDu, (unsigned int)(g >> 32) << 1, p.seed_lo, p.seed_hi), one, a);
        philox_normal4(philox4x32((unsigned int)g, p.tick, 0x00C105EDu, ((unsigned int)(g >> 32) << 1) | 1u, p.seed_lo, p.seed_hi), one, b);
#pragma unroll
        for (int i = 0; i < 4; ++i) { n[i] = a[i]; n[4 + i] = b[i]; }
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) p.z[(long long)i * p.C + c] = z[i] + p.r_std[i] * (double)n[i];
}

}  // namespace mpcb

using namespace mpcb;

struct mpcb_closed_loop {
    mpcb_closed_loop_cfg cfg;
    mpcb_mppi* mppi = nullptr;
    mpcb_ukf* ukf = nullptr;
    long long C = 0;
    int H = 0;
    double* d_x = nullptr;   // [6][C]
    double* d_z = nullptr;   // [5][C]
    double* d_x4 = nullptr;  // [C][4]
    double* d_u[2] = {nullptr, nullptr};  // [C][H] ping-pong
    double* d_u0 = nullptr;  // [C]
    void* d_eps = nullptr;
    size_t d_eps_bytes = 0;
    int cur = 0;
    long long ticks = 0;
    cudaEvent_t ev_ukf = nullptr, ev_mppi = nullptr;
    PlantParams pp;
    int64_t launches = 0;
};

namespace mpcb {
// [C][4] <- rows (0, 1, 3, 4) of the truth (SoA [6][C])
__global__ void closed_loop_gather_truth_kernel(const double* x, double* x4, long long C) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    x4[c * 4 + 0] = x[c];
    x4[c * 4 + 1] = x[C + c];
    x4[c * 4 + 2] = x[3 * C + c];
    x4[c * 4 + 3] = x[4 * C + c];
}
}  // namespace mpcb

namespace {
mpcb_status gather_truth(mpcb_closed_loop* h) {
    closed_loop_gather_truth_kernel<<<(unsigned)((h->C + 127) / 128), 128, 0, (cudaStream_t)mpcb_ukf_stream(h->ukf)>>>(h->d_x, h->d_x4, h->C);
    MPCB_CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return MPCB_OK;
}

mpcb_status tick_once(mpcb_closed_loop* h, const double* z_host, const void* eps_host, int eps_dtype) {
    cudaStream_t su = (cudaStream_t)mpcb_ukf_stream(h->ukf), sm = (cudaStream_t)mpcb_mppi_stream(h->mppi);
    const double t = (double)h->ticks * h->cfg.tick_dt;
    // plant + sensor (on the UKF's stream, after the previous tick's MPPI produced u_n[0])
    MPCB_CUDA_TRY(cudaStreamWaitEvent(su, h->ev_mppi, 0));
    PlantParams p = h->pp;
    p.f = (t > 1.0 && t < 1.5) ? 2.0 : 0.0;  // the 2 N push of :236-241
    p.tick = (unsigned int)h->ticks;
    p.write_z = z_host ? 0 : 1;
    closed_loop_plant_kernel<<<(unsigned)((h->C + 127) / 128), 128, 0, su>>>(p);
    MPCB_CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    if (z_host) {  // replay: the caller's readings [C][5] -> SoA [5][C]
        // (verification path: a synchronous staging copy is fine)
        double* tmp = new (std::nothrow) double[(size_t)5 * h->C];
        if (!tmp) { set_error("out of host memory"); return MPCB_BAD_ARG; }
        for (long long c = 0; c < h->C; ++c)
            for (int i = 0; i < 5; ++i) tmp[(size_t)i * h->C + c] = z_host[(size_t)c * 5 + i];
        cudaError_t e = cudaMemcpyAsync(h->d_z, tmp, sizeof(double) * 5 * h->C, cudaMemcpyHostToDevice, su);
        if (e == cudaSuccess) e = cudaStreamSynchronize(su);
        delete[] tmp;
        MPCB_CUDA_TRY(e);
    }
    // UKF: fused predict(u_n[0]) + update(z) with the control being applied, then the estimate's [x0, x1, x3, x4] (:78)
    mpcb_status st = mpcb_ukf_run_device(h->ukf, 1, h->d_u0, 0.0, h->cfg.tick_dt, h->d_z);
    if (st != MPCB_OK) return st;
    if (h->cfg.use_estimate) {
        const int32_t idx[4] = {0, 1, 3, 4};
        st = mpcb_ukf_gather_state_device(h->ukf, 4, idx, h->d_x4);
        if (st != MPCB_OK) return st;
    } else {
        // DEBUG_UKF = true (:55-57): the controller sees the true state
        st = gather_truth(h);
        if (st != MPCB_OK) return st;
    }
    MPCB_CUDA_TRY(cudaEventRecord(h->ev_ukf, su));
    // MPPI: all C controllers in one launch, previous sequence in, new sequence out; u_n[0] for the next tick
    MPCB_CUDA_TRY(cudaStreamWaitEvent(sm, h->ev_ukf, 0));
    const void* d_eps = nullptr;
    if (eps_host) {
        const size_t es = eps_dtype == MPCB_DT_F64 ? 8 : 4;
        const size_t bytes = (size_t)h->C * (size_t)h->cfg.samples * h->H * es;
        if (h->d_eps_bytes < bytes) {
            if (h->d_eps) cudaFree(h->d_eps);
            h->d_eps = nullptr;
            h->d_eps_bytes = 0;
            MPCB_CUDA_TRY(cudaMalloc(&h->d_eps, bytes));
            h->d_eps_bytes = bytes;
        }
        MPCB_CUDA_TRY(cudaMemcpyAsync(h->d_eps, eps_host, bytes, cudaMemcpyHostToDevice, sm));
        d_eps = h->d_eps;
    }
    const int nxt = h->cur ^ 1;
    st = mpcb_mppi_compute_device(h->mppi, h->d_x4, h->d_u[h->cur], d_eps, eps_dtype, h->d_u[nxt]);
    if (st != MPCB_OK) return st;
    st = mpcb_mppi_first_control_device(h->mppi, h->d_u[nxt], h->d_u0);
    if (st != MPCB_OK) return st;
    MPCB_CUDA_TRY(cudaEventRecord(h->ev_mppi, sm));
    h->cur = nxt;
    h->ticks += 1;
    return MPCB_OK;
}
}  // namespace

extern "C" {

mpcb_status mpcb_closed_loop_default_cfg(mpcb_closed_loop_cfg* out) {
    MPCB_REQUIRE(out != nullptr, "null pointer");
    memset(out, 0, sizeof(*out));
    out->controllers = 4096;
    out->samples = 8192;
    out->tick_dt = 0.01;  // the reference's ~9-10 ms sensor period (:267-268) as a fixed tick
    out->use_estimate = 1;
    out->precision = -1;  // the MPPI default for model NL6
    out->exact_ukf = 0;
    out->seed = 20240004;
    out->device = 0;
    out->controller_offset = 0;
    return MPCB_OK;
}

mpcb_status mpcb_closed_loop_create(mpcb_closed_loop** out, const mpcb_closed_loop_cfg* cfg) {
    MPCB_REQUIRE(out && cfg, "null pointer");
    MPCB_REQUIRE(cfg->controllers >= 1 && cfg->samples >= 1 && cfg->tick_dt > 0.0 && cfg->controller_offset >= 0, "bad closed-loop shape");
    *out = nullptr;
    mpcb_closed_loop* h = new (std::nothrow) mpcb_closed_loop;
    MPCB_REQUIRE(h != nullptr, "out of memory");
    h->cfg = *cfg;
    h->C = cfg->controllers;
    auto fail = [&](mpcb_status st) {
        mpcb_closed_loop_destroy(h);
        return st;
    };
    // MPPI: the constants of examples/mppi4-non-liner-ukf.rs:13-24 (T = 1.2, N = 8, lambda = 1.4, R = 4, limit +-10)
    mpcb_mppi_cfg mc;
    mpcb_status st = mpcb_mppi_default_cfg(MPCB_MODEL_NL6, &mc);
    if (st != MPCB_OK) return fail(st);
    mc.samples = cfg->samples;
    mc.controllers = cfg->controllers;
    mc.seed = cfg->seed;
    mc.device = cfg->device;
    if (cfg->precision == MPCB_F32 || cfg->precision == MPCB_F64 || cfg->precision == MPCB_F64_FAST) mc.precision = cfg->precision;
    h->H = mc.horizon;
    st = mpcb_mppi_create(&h->mppi, &mc);
    if (st != MPCB_OK) return fail(st);
    st = mpcb_mppi_set_controller_offset(h->mppi, cfg->controller_offset);
    if (st != MPCB_OK) return fail(st);
    // UKF: model NL6_UKF, one filter per robot, Q = gen_q(tick_dt) (:192-221), R (:28), P0 = 10 I (:163)
    mpcb_ukf_cfg uc;
    st = mpcb_ukf_default_cfg(MPCB_MODEL_NL6_UKF, &uc);
    if (st != MPCB_OK) return fail(st);
    uc.batch = cfg->controllers;
    uc.device = cfg->device;
    uc.exact = cfg->exact_ukf;
    st = mpcb_ukf_create(&h->ukf, &uc);
    if (st != MPCB_OK) return fail(st);
    double Q[36], R[25], P0[36], x0[6] = {0, 0, 0, 0, 0, 0};
    st = mpcb_ukf_default_noise(MPCB_MODEL_NL6_UKF, cfg->tick_dt, Q, R, P0);
    if (st != MPCB_OK) return fail(st);
    st = mpcb_ukf_init(h->ukf, x0, P0, Q, R);
    if (st != MPCB_OK) return fail(st);
#define TRY_OR_FAIL(expr)                                                              \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                 \
            return fail(MPCB_CUDA_ERROR);                                              \
        }                                                                              \
    } while (0)
    TRY_OR_FAIL(cudaSetDevice(cfg->device));
    const size_t C = (size_t)h->C, H = (size_t)h->H;
    TRY_OR_FAIL(cudaMalloc(&h->d_x, 6 * C * sizeof(double)));
    TRY_OR_FAIL(cudaMalloc(&h->d_z, 5 * C * sizeof(double)));
    TRY_OR_FAIL(cudaMalloc(&h->d_x4, 4 * C * sizeof(double)));
    TRY_OR_FAIL(cudaMalloc(&h->d_u[0], C * H * sizeof(double)));
    TRY_OR_FAIL(cudaMalloc(&h->d_u[1], C * H * sizeof(double)));
    TRY_OR_FAIL(cudaMalloc(&h->d_u0, C * sizeof(double)));
    TRY_OR_FAIL(cudaMemset(h->d_x, 0, 6 * C * sizeof(double)));
    TRY_OR_FAIL(cudaMemset(h->d_z, 0, 5 * C * sizeof(double)));
    TRY_OR_FAIL(cudaMemset(h->d_u[0], 0, C * H * sizeof(double)));
    TRY_OR_FAIL(cudaMemset(h->d_u[1], 0, C * H * sizeof(double)));
    TRY_OR_FAIL(cudaMemset(h->d_u0, 0, C * sizeof(double)));
    TRY_OR_FAIL(cudaDeviceSynchronize());
    TRY_OR_FAIL(cudaEventCreateWithFlags(&h->ev_ukf, cudaEventDisableTiming));
    TRY_OR_FAIL(cudaEventCreateWithFlags(&h->ev_mppi, cudaEventDisableTiming));
    TRY_OR_FAIL(cudaEventRecord(h->ev_mppi, (cudaStream_t)mpcb_mppi_stream(h->mppi)));
#undef TRY_OR_FAIL
    mpcb_model_params mp;
    st = mpcb_model_defaults(MPCB_MODEL_NL6, &mp);
    if (st != MPCB_OK) return fail(st);
    PlantParams& p = h->pp;
    memset(&p, 0, sizeof(p));
    p.x = h->d_x; p.u0 = h->d_u0; p.z = h->d_z;
    p.C = h->C; p.c_offset = cfg->controller_offset;
    p.dt = cfg->tick_dt;
    p.seed_lo = (unsigned int)(cfg->seed & 0xffffffffull);
    p.seed_hi = (unsigned int)(cfg->seed >> 32);
    p.M1 = mp.m1; p.R_W = mp.r_w; p.M2 = mp.m2; p.L = mp.l; p.J1 = mp.j1; p.J2 = mp.j2; p.G = mp.g; p.KT = mp.kt;
    const double rs[5] = {200.0, 200.0, 10.0, 0.05, 0.05};  // :28, used as standard deviations by the sensor (:186)
    for (int i = 0; i < 5; ++i) p.r_std[i] = rs[i];
    *out = h;
    return MPCB_OK;
}

void mpcb_closed_loop_destroy(mpcb_closed_loop* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->mppi) mpcb_mppi_sync(h->mppi);
    if (h->ukf) mpcb_ukf_sync(h->ukf);
    cudaFree(h->d_x);
    cudaFree(h->d_z);
    cudaFree(h->d_x4);
    cudaFree(h->d_u[0]);
    cudaFree(h->d_u[1]);
    cudaFree(h->d_u0);
    cudaFree(h->d_eps);
    if (h->ev_ukf) cudaEventDestroy(h->ev_ukf);
    if (h->ev_mppi) cudaEventDestroy(h->ev_mppi);
    if (h->mppi) mpcb_mppi_destroy(h->mppi);
    if (h->ukf) mpcb_ukf_destroy(h->ukf);
    cudaGetLastError();
    delete h;
}

mpcb_status mpcb_closed_loop_set_state(mpcb_closed_loop* h, const double* x6, const double* P) {
    MPCB_REQUIRE(h && x6, "null pointer");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    mpcb_status st = mpcb_closed_loop_sync(h);
    if (st != MPCB_OK) return st;
    // truth (SoA) and the filters' estimate: init_ukf(&init_x) (:40,161-167)
    double* tmp = new (std::nothrow) double[(size_t)6 * h->C];
    MPCB_REQUIRE(tmp != nullptr, "out of host memory");
    for (long long c = 0; c < h->C; ++c)
        for (int i = 0; i < 6; ++i) tmp[(size_t)i * h->C + c] = x6[(size_t)c * 6 + i];
    cudaError_t e = cudaMemcpy(h->d_x, tmp, sizeof(double) * 6 * h->C, cudaMemcpyHostToDevice);
    delete[] tmp;
    MPCB_CUDA_TRY(e);
    return mpcb_ukf_set_state(h->ukf, x6, P);
}

mpcb_status mpcb_closed_loop_set_truth(mpcb_closed_loop* h, const double* x6) {
    MPCB_REQUIRE(h && x6, "null pointer");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    mpcb_status st = mpcb_closed_loop_sync(h);
    if (st != MPCB_OK) return st;
    double* tmp = new (std::nothrow) double[(size_t)6 * h->C];
    MPCB_REQUIRE(tmp != nullptr, "out of host memory");
    for (long long c = 0; c < h->C; ++c)
        for (int i = 0; i < 6; ++i) tmp[(size_t)i * h->C + c] = x6[(size_t)c * 6 + i];
    cudaError_t e = cudaMemcpy(h->d_x, tmp, sizeof(double) * 6 * h->C, cudaMemcpyHostToDevice);
    delete[] tmp;
    MPCB_CUDA_TRY(e);
    return MPCB_OK;
}

mpcb_status mpcb_closed_loop_set_controls(mpcb_closed_loop* h, const double* u_seq) {
    MPCB_REQUIRE(h && u_seq, "null pointer");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    mpcb_status st = mpcb_closed_loop_sync(h);
    if (st != MPCB_OK) return st;
    MPCB_CUDA_TRY(cudaMemcpy(h->d_u[h->cur], u_seq, sizeof(double) * h->C * h->H, cudaMemcpyHostToDevice));
    return mpcb_mppi_first_control_device(h->mppi, h->d_u[h->cur], h->d_u0) == MPCB_OK ? mpcb_mppi_sync(h->mppi) : MPCB_CUDA_ERROR;
}

mpcb_status mpcb_closed_loop_tick(mpcb_closed_loop* h, int32_t n_ticks) {
    MPCB_REQUIRE(h && n_ticks >= 0, "bad arguments");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    for (int i = 0; i < n_ticks; ++i) {
        mpcb_status st = tick_once(h, nullptr, nullptr, MPCB_DT_F32);
        if (st != MPCB_OK) return st;
    }
    return MPCB_OK;
}

mpcb_status mpcb_closed_loop_tick_replay(mpcb_closed_loop* h, const double* z, const void* eps, int32_t eps_dtype) {
    MPCB_REQUIRE(h != nullptr, "null pointer");
    MPCB_REQUIRE(eps == nullptr || eps_dtype == MPCB_DT_F32 || eps_dtype == MPCB_DT_F64, "bad eps dtype");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    mpcb_status st = tick_once(h, z, eps, eps ? eps_dtype : MPCB_DT_F32);
    if (st != MPCB_OK) return st;
    return mpcb_closed_loop_sync(h);  // the caller's eps may be reused right away
}

mpcb_status mpcb_closed_loop_sync(mpcb_closed_loop* h) {
    MPCB_REQUIRE(h != nullptr, "null pointer");
    mpcb_status st = mpcb_ukf_sync(h->ukf);
    if (st != MPCB_OK) return st;
    return mpcb_mppi_sync(h->mppi);
}

mpcb_status mpcb_closed_loop_get(mpcb_closed_loop* h, double* x6, double* x_est, double* z, double* u0, double* u_seq,
                                 int32_t* mppi_status) {
    MPCB_REQUIRE(h != nullptr, "null pointer");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    mpcb_status st = mpcb_closed_loop_sync(h);
    if (st != MPCB_OK) return st;
    const size_t C = (size_t)h->C;
    if (x6 || z) {
        double* tmp = new (std::nothrow) double[6 * C];
        MPCB_REQUIRE(tmp != nullptr, "out of host memory");
        cudaError_t e = cudaSuccess;
        if (x6) {
            e = cudaMemcpy(tmp, h->d_x, sizeof(double) * 6 * C, cudaMemcpyDeviceToHost);
            for (size_t c = 0; c < C && e == cudaSuccess; ++c)
                for (int i = 0; i < 6; ++i) x6[c * 6 + i] = tmp[(size_t)i * C + c];
        }
        if (z && e == cudaSuccess) {
            e = cudaMemcpy(tmp, h->d_z, sizeof(double) * 5 * C, cudaMemcpyDeviceToHost);
            for (size_t c = 0; c < C && e == cudaSuccess; ++c)
                for (int i = 0; i < 5; ++i) z[c * 5 + i] = tmp[(size_t)i * C + c];
        }
        delete[] tmp;
        MPCB_CUDA_TRY(e);
    }
    if (x_est) {
        st = mpcb_ukf_get_state(h->ukf, x_est, nullptr);
        if (st != MPCB_OK) return st;
    }
    if (u0) MPCB_CUDA_TRY(cudaMemcpy(u0, h->d_u0, sizeof(double) * C, cudaMemcpyDeviceToHost));
    if (u_seq) MPCB_CUDA_TRY(cudaMemcpy(u_seq, h->d_u[h->cur], sizeof(double) * C * h->H, cudaMemcpyDeviceToHost));
    if (mppi_status) {
        mpcb_mppi_info* infos = new (std::nothrow) mpcb_mppi_info[C];
        MPCB_REQUIRE(infos != nullptr, "out of host memory");
        st = mpcb_mppi_last_info(h->mppi, infos);
        for (size_t c = 0; c < C; ++c) mppi_status[c] = infos[c].status;
        delete[] infos;
        if (st != MPCB_OK) return st;
    }
    return MPCB_OK;
}

mpcb_mppi* mpcb_closed_loop_mppi(mpcb_closed_loop* h) { return h ? h->mppi : nullptr; }
mpcb_ukf* mpcb_closed_loop_ukf(mpcb_closed_loop* h) { return h ? h->ukf : nullptr; }
int64_t mpcb_closed_loop_ticks(mpcb_closed_loop* h) { return h ? h->ticks : 0; }
int64_t mpcb_closed_loop_launches(mpcb_closed_loop* h) {
    return h ? h->launches + mpcb_mppi_launches(h->mppi) + mpcb_ukf_launches(h->ukf) : 0;
}

}  // extern "C"
