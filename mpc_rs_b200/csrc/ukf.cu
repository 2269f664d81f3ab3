// ukf.cu — C ABI of the batched UKF (mpcb_ukf_*), replacing mpc::ukf / mpc::ukf2::UnscentedKalmanFilter
// (src/ukf.rs:30-94, src/ukf2.rs:30-98) and the free functions of examples/ukf-pen.rs:93-141.
#include <math.h>

#include <new>

#include <string>

#include "mppi_rtc.h"
#include "ukf_kernel.cuh"

namespace mpcb {

// in[B][W] (array of structures, host-facing) -> out[W][B] (structure of arrays, device layout)
__global__ void ukf_aos_to_soa(const double* __restrict__ in, double* __restrict__ out, long long B, int W) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // index into out: w*B + b
    if (i >= B * W) return;
    const long long w = i / B, b = i - w * B;
    out[i] = in[b * W + w];
}
// in[W][B] -> out[count][W] for filters [first, first+count)
__global__ void ukf_soa_to_aos(const double* __restrict__ in, double* __restrict__ out, long long B, int W,
                               long long first, long long count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // index over w*count + j (coalesced reads)
    if (i >= count * W) return;
    const long long w = i / count, j = i - w * count;
    out[j * W + w] = in[w * B + first + j];
}
__global__ void ukf_broadcast(const double* __restrict__ row, double* __restrict__ out, long long B, int W) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * W) return;
    out[i] = row[i / B];
}

}  // namespace mpcb

using namespace mpcb;

struct mpcb_ukf {
    mpcb_ukf_cfg cfg;
    int n = 0, o = 0, M = 0;
    long long B = 0;
    cudaStream_t stream = nullptr;
    ModelConsts mc;
    UkfKernelFn k_predict = nullptr, k_update = nullptr, k_fused = nullptr;
    size_t fused_smem = 0;  // dynamic shared memory of k_fused (the streaming six-state kernels keep the sigma points there)
    long long grid_cap = 0;  // resident blocks of the fused kernel on this device
    RtcModule rtc;  // user-supplied fx / hx (mpcb_ukf_create_user): the kernels above live in this module
    double Q[36], R[25];
    double wm0 = 0, wc0 = 0, wi = 0, cC = 0;
    double* d_x = nullptr;
    double* d_P = nullptr;
    double* d_sigma = nullptr;
    double* d_stage = nullptr;  // staging for host AoS transfers: max(n*n, o) doubles per filter
    double* d_u = nullptr;
    double* d_z = nullptr;
    int* d_status = nullptr;
    bool predicted = false;
    bool p_lower_only = false;  // the last fused step stored only the lower triangle of P: read-outs mirror it
    unsigned int enable = 0xffffffffu;  // sensor mask of the next update (examples/mppi4-ukf-commu.rs:279-293)
    int64_t launches = 0;
};

// x[B][n_idx] (AoS rows) <- the state components idx[] of every filter (SoA [n][B])
struct GatherParams {
    const double* x;
    double* out;
    long long B;
    int n_idx;
    int idx[8];
};
__global__ void ukf_gather_state_kernel(const GatherParams p) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    for (int j = 0; j < p.n_idx; ++j) p.out[b * p.n_idx + j] = p.x[(long long)p.idx[j] * p.B + b];
}

namespace {

constexpr int kThreads = 128;
inline unsigned blocks_for(long long n) { return (unsigned)((n + 255) / 256); }

mpcb_status model_dims(int model_id, int* n, int* o) {
    switch (model_id) {
        case MPCB_MODEL_PEN_LIN: *n = 4; *o = 2; return MPCB_OK;
        case MPCB_MODEL_PEN_NL: *n = 4; *o = 3; return MPCB_OK;
        case MPCB_MODEL_PEN6:
        case MPCB_MODEL_NL6_UKF: *n = 6; *o = 5; return MPCB_OK;
        default: set_error("model %d is not a UKF model", model_id); return MPCB_BAD_ARG;
    }
}

// examples/mppi4-non-liner-ukf.rs:192-221, PHY = (100, 70, 20) (:27)
void gen_q(double dt, double* q) {
    const double dt_2 = dt * dt, dt_3 = dt_2 * dt, dt_4 = dt_2 * dt_2;
    double q1[36] = {0}, q2[36] = {0}, q3[36] = {0};
    auto at = [](double* m, int r, int c) -> double& { return m[r * 6 + c]; };
    at(q1, 3, 4) = dt_4 / 8.0; at(q1, 3, 5) = dt_3 / 6.0;
    at(q1, 4, 3) = dt_4 / 8.0; at(q1, 4, 4) = dt_3 / 3.0; at(q1, 4, 5) = dt_2 / 2.0;
    at(q1, 5, 3) = dt_3 / 6.0; at(q1, 5, 4) = dt_2 / 2.0; at(q1, 5, 5) = dt;
    at(q2, 1, 3) = dt_4 / 8.0; at(q2, 1, 4) = dt_3 / 6.0;
    at(q2, 3, 1) = dt_4 / 8.0; at(q2, 3, 3) = dt_3 / 3.0; at(q2, 3, 4) = dt_2 / 2.0;
    at(q2, 4, 1) = dt_3 / 6.0; at(q2, 4, 3) = dt_2 / 2.0; at(q2, 4, 4) = dt;
    at(q3, 0, 1) = dt_4 / 8.0; at(q3, 0, 2) = dt_3 / 6.0;
    at(q3, 1, 0) = dt_4 / 8.0; at(q3, 1, 1) = dt_3 / 3.0; at(q3, 1, 2) = dt_2 / 2.0;
    at(q3, 2, 0) = dt_3 / 6.0; at(q3, 2, 1) = dt_2 / 2.0; at(q3, 2, 2) = dt;
    for (int i = 0; i < 36; ++i) q[i] = 100.0 * q1[i] + 70.0 * q2[i] + 20.0 * q3[i];
}

void fill_params(const mpcb_ukf* h, UkfParams* p) {
    memset(p, 0, sizeof(*p));
    p->B = h->B;
    p->steps = 1;
    p->x = h->d_x;
    p->P = h->d_P;
    p->sigma_f = h->d_sigma;
    p->status = h->d_status;
    p->dt = h->cfg.model.dt;
    p->wm0 = h->wm0; p->wc0 = h->wc0; p->wi = h->wi; p->cC = h->cC;
    memcpy(p->Q, h->Q, sizeof(p->Q));
    memcpy(p->R, h->R, sizeof(p->R));
    p->mc = h->mc;
    p->enable = h->enable;
    p->reverse = (getenv("MPCB_UKF_NO_REVERSE") == nullptr) ? (unsigned int)(h->launches & 1) : 0u;
}

mpcb_status launch(mpcb_ukf* h, UkfKernelFn fn, const UkfParams& p) {
    // the kernels walk tiles of 128 filters grid-stride; at most one resident wave of blocks, so that the fused kernel's
    // prefetch of the next tile overlaps the current tile's arithmetic
    long long tiles = (h->B + kThreads - 1) / kThreads;
    if (h->grid_cap > 0 && tiles > h->grid_cap) tiles = h->grid_cap;
    const unsigned grid = (unsigned)tiles;
    // explicit cudaLaunchKernel: `fn` is a compiled-in __global__ function or the cudaKernel_t of a user model
    UkfParams pp = p;
    void* args[1] = {&pp};
    const size_t smem = (fn == h->k_fused) ? h->fused_smem : 0;
    MPCB_CUDA_TRY(launch_pdl(reinterpret_cast<const void*>(fn), dim3(grid), dim3(kThreads), args, smem, h->stream));
    h->launches += 1;
    return MPCB_OK;
}

// host AoS rows [B][W] -> device SoA [W][B] through the staging buffer
mpcb_status upload_aos(mpcb_ukf* h, const double* host, int W, double* d_soa) {
    const size_t bytes = (size_t)h->B * W * sizeof(double);
    MPCB_CUDA_TRY(cudaMemcpyAsync(h->d_stage, host, bytes, cudaMemcpyHostToDevice, h->stream));
    ukf_aos_to_soa<<<blocks_for(h->B * W), 256, 0, h->stream>>>(h->d_stage, d_soa, h->B, W);
    MPCB_CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return MPCB_OK;
}

mpcb_status ensure_sigma(mpcb_ukf* h) {
    if (h->d_sigma) return MPCB_OK;
    MPCB_CUDA_TRY(cudaMalloc(&h->d_sigma, (size_t)h->B * h->n * h->M * sizeof(double)));
    return MPCB_OK;
}
mpcb_status ensure_u(mpcb_ukf* h) {
    if (h->d_u) return MPCB_OK;
    MPCB_CUDA_TRY(cudaMalloc(&h->d_u, (size_t)h->B * sizeof(double)));
    return MPCB_OK;
}
mpcb_status ensure_z(mpcb_ukf* h) {
    if (h->d_z) return MPCB_OK;
    MPCB_CUDA_TRY(cudaMalloc(&h->d_z, (size_t)h->B * h->o * sizeof(double)));
    return MPCB_OK;
}

mpcb_status stage_u(mpcb_ukf* h, const double* u, double u_scalar, UkfParams* p) {
    if (u) {
        mpcb_status st = ensure_u(h);
        if (st != MPCB_OK) return st;
        MPCB_CUDA_TRY(cudaMemcpyAsync(h->d_u, u, (size_t)h->B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        p->u = h->d_u;
        p->has_u = 1;
    } else {
        p->u_scalar = u_scalar;
        p->has_u = 0;
    }
    return MPCB_OK;
}

mpcb_status stage_z(mpcb_ukf* h, const double* z, UkfParams* p) {
    mpcb_status st = ensure_z(h);
    if (st != MPCB_OK) return st;
    st = upload_aos(h, z, h->o, h->d_z);
    if (st != MPCB_OK) return st;
    p->z = h->d_z;
    return MPCB_OK;
}

}  // namespace

extern "C" {

mpcb_status mpcb_ukf_default_cfg(int32_t model_id, mpcb_ukf_cfg* c) {
    if (!c) return MPCB_BAD_ARG;
    memset(c, 0, sizeof(*c));
    if (model_id == MPCB_MODEL_USER_UKF) {  // the library UKF's defaults; n, o and the model come from the caller
        c->model_id = model_id;
        c->sqrt_mode = MPCB_SQRT_EIG;
        c->sigma_order = MPCB_ORDER_LIBRARY;
        c->batch = 1;
        return MPCB_OK;
    }
    int n, o;
    mpcb_status st = model_dims(model_id, &n, &o);
    if (st != MPCB_OK) return st;
    st = mpcb_model_defaults(model_id, &c->model);
    if (st != MPCB_OK) return st;
    c->model_id = model_id;
    c->n = n;
    c->o = o;
    // examples/ukf-pen.rs uses Cholesky + interleaved columns; mpc::ukf / mpc::ukf2 use the SVD square root
    c->sqrt_mode = (model_id == MPCB_MODEL_PEN_LIN) ? MPCB_SQRT_CHOLESKY : MPCB_SQRT_EIG;
    c->sigma_order = (model_id == MPCB_MODEL_PEN_LIN) ? MPCB_ORDER_INTERLEAVED : MPCB_ORDER_LIBRARY;
    c->batch = 1;
    return MPCB_OK;
}

mpcb_status mpcb_ukf_default_noise(int32_t model_id, double dt, double* Q, double* R, double* P0) {
    MPCB_REQUIRE(Q && R && P0, "null pointer");
    int n, o;
    mpcb_status st = model_dims(model_id, &n, &o);
    if (st != MPCB_OK) return st;
    memset(Q, 0, sizeof(double) * n * n);
    memset(R, 0, sizeof(double) * o * o);
    memset(P0, 0, sizeof(double) * n * n);
    for (int i = 0; i < n; ++i) P0[i * n + i] = 10.0;
    switch (model_id) {
        case MPCB_MODEL_PEN_LIN:  // examples/ukf-pen.rs:17-26,148-153
            Q[1 * 4 + 1] = 1.0; Q[2 * 4 + 2] = 0.25; Q[2 * 4 + 3] = 0.5; Q[3 * 4 + 2] = 0.5; Q[3 * 4 + 3] = 1.0;
            R[0] = 0.5; R[3] = 0.5;
            break;
        case MPCB_MODEL_PEN_NL:  // examples/ukf-pen2.rs:18-28,71-76
            Q[3 * 4 + 3] = 0.25;
            R[0] = 100.0; R[4] = 100.0; R[8] = 0.5;
            break;
        case MPCB_MODEL_PEN6:  // examples/ukf-pen3.rs:18-32,83-90
            Q[5 * 6 + 5] = 10.0;
            R[0] = 100.0; R[6] = 100.0; R[12] = 0.5; R[18] = 100.0; R[24] = 100.0;
            break;
        default: {  // MPCB_MODEL_NL6_UKF: examples/mppi4-non-liner-ukf.rs:28,161-167
            const double rdiag[5] = {200.0, 200.0, 10.0, 0.05, 0.05};
            gen_q(dt > 0.0 ? dt : 1.2 / 8.0, Q);
            for (int i = 0; i < 5; ++i) R[i * 5 + i] = rdiag[i];
            break;
        }
    }
    return MPCB_OK;
}

static mpcb_status ukf_create_impl(mpcb_ukf** out, const mpcb_ukf_cfg* cfg, const char* user_src, const double* params,
                                   int32_t n_params) {
    MPCB_REQUIRE(out && cfg, "null pointer");
    *out = nullptr;
    int n, o;
    mpcb_status st = MPCB_OK;
    if (user_src != nullptr) {
        n = cfg->n;
        o = cfg->o;
        MPCB_REQUIRE(n >= 1 && n <= 6 && o >= 1 && o <= 5, "user UKF models: n in 1..6, o in 1..5");
    } else {
        st = model_dims(cfg->model_id, &n, &o);
        if (st != MPCB_OK) return st;
        MPCB_REQUIRE(cfg->n == n && cfg->o == o, "n/o do not match the model");
    }
    MPCB_REQUIRE(cfg->batch >= 1, "batch must be >= 1");
    MPCB_REQUIRE(cfg->sqrt_mode == MPCB_SQRT_CHOLESKY || cfg->sqrt_mode == MPCB_SQRT_EIG, "bad sqrt_mode");
    MPCB_REQUIRE(cfg->sigma_order == MPCB_ORDER_LIBRARY || cfg->sigma_order == MPCB_ORDER_INTERLEAVED, "bad sigma_order");
    MPCB_REQUIRE(user_src != nullptr || !(n == 6 && cfg->sigma_order != MPCB_ORDER_LIBRARY),
                 "n = 6 supports the library sigma order only");
    int ndev = 0;
    MPCB_CUDA_TRY(cudaGetDeviceCount(&ndev));
    MPCB_REQUIRE(cfg->device >= 0 && cfg->device < ndev, "no such CUDA device (this library has no CPU path)");
    MPCB_CUDA_TRY(cudaSetDevice(cfg->device));

    mpcb_ukf* h = new (std::nothrow) mpcb_ukf();
    MPCB_REQUIRE(h != nullptr, "out of memory");
    h->cfg = *cfg;
    h->n = n;
    h->o = o;
    h->M = 2 * n + 1;
    h->B = cfg->batch;
    auto fail = [&](mpcb_status s) {
        mpcb_ukf_destroy(h);
        return s;
    };
    if (user_src != nullptr) {
        h->cfg.model_id = MPCB_MODEL_USER_UKF;
        memset(&h->mc, 0, sizeof(h->mc));
        for (int i = 0; i < n_params; ++i) {
            h->mc.k[i] = params[i];
            h->mc.kf[i] = (float)params[i];
        }
    } else {
        st = build_model_consts(cfg->model_id, cfg->model, cfg->model.dt, &h->mc);
        if (st != MPCB_OK) return fail(st);
    }
    if (user_src == nullptr) {
        // measurement constants, evaluated like the reference's expressions
        const double PI = 3.14159265358979323846264338327950288;
        const mpcb_model_params& mp = cfg->model;
        double* k = h->mc.k;
        k[uslot::G] = mp.g;
        k[uslot::IG] = 1.0 / mp.g;
        k[uslot::L] = mp.l;
        k[uslot::DEG] = 180.0 / PI;  // f64::to_degrees
        k[uslot::M2G] = mp.m2 * mp.g;
        k[uslot::M2] = mp.m2;
        k[uslot::M2L] = mp.m2 * mp.l;
        if (cfg->model_id == MPCB_MODEL_NL6_UKF) {
            k[uslot::RPM] = 36.0 * 60.0 / (2.0 * PI * mp.r_w);    // examples/mppi4-non-liner-ukf.rs:173
            k[uslot::NRPM] = 36.0 * -60.0 / (2.0 * PI * mp.r_w);  // :174
        } else {
            k[uslot::RPM] = 60.0 / (2.0 * PI * mp.r_w);  // examples/ukf-pen2.rs:49
            k[uslot::NRPM] = -k[uslot::RPM];
        }
    }
    {
        // sigma_weight, src/ukf.rs:23-28,112-118
        const double N = (double)n, ALPHA = 1e-3, BETA = 2.0;
        const double KAPPA = 3.0 - N;
        const double Cc = ALPHA * ALPHA * (N + KAPPA);
        const double LAMBDA = Cc - N;
        h->cC = Cc;
        h->wi = 1.0 / (2.0 * Cc);
        h->wm0 = LAMBDA / Cc;
        h->wc0 = LAMBDA / Cc + 1.0 - ALPHA * ALPHA + BETA;
    }
    if (user_src != nullptr) {
        st = rtc_compile_ukf_user(user_src, n, o, cfg->sqrt_mode, cfg->sigma_order, cfg->exact == 0, true, &h->rtc);
        if (st != MPCB_OK) return fail(st);
        h->k_predict = reinterpret_cast<UkfKernelFn>(h->rtc.kernel[UKF_PREDICT]);
        h->k_update = reinterpret_cast<UkfKernelFn>(h->rtc.kernel[UKF_UPDATE]);
        h->k_fused = reinterpret_cast<UkfKernelFn>(h->rtc.kernel[UKF_FUSED]);
    } else {
        auto pick = cfg->exact ? ((n == 4) ? ukf_kernel_n4 : ukf_kernel_n6) : ((n == 4) ? ukf_kernel_n4_fast : ukf_kernel_n6_fast);
        h->k_predict = pick(cfg->model_id, cfg->sqrt_mode, cfg->sigma_order, UKF_PREDICT);
        h->k_update = pick(cfg->model_id, cfg->sqrt_mode, cfg->sigma_order, UKF_UPDATE);
        h->k_fused = pick(cfg->model_id, cfg->sqrt_mode, cfg->sigma_order, UKF_FUSED);
        // six-state filters in fast arithmetic: the split predict / update calls run the reference-order kernels — the general
        // fast kernel's unshifted sums lose the state on NL6_UKF with covariances >> 1 (tools/dev_ukf6_random.py: Cholesky,
        // 8 x a Wishart matrix, x off by 7 where the reference-order and the streaming kernel agree with the oracle to 3e-5) —
        // and the fused step runs the streaming kernel (ukf_stream_kernel.cuh: sigma
        // points in shared memory, one-pass shifted transforms); MPCB_UKF_NO_STREAM=1 keeps the general kernel (A/B)
        if (!cfg->exact && n == 6 && !getenv("MPCB_UKF_NO_STREAM")) {
            h->k_predict = ukf_kernel_n6(cfg->model_id, cfg->sqrt_mode, cfg->sigma_order, UKF_PREDICT);
            h->k_update = ukf_kernel_n6(cfg->model_id, cfg->sqrt_mode, cfg->sigma_order, UKF_UPDATE);
            size_t smem = 0;
            UkfKernelFn ks = ukf_stream_kernel_n6(cfg->model_id, cfg->sqrt_mode, cfg->sigma_order, &smem);
            if (ks != nullptr &&
                cudaFuncSetAttribute(reinterpret_cast<const void*>(ks), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
                h->k_fused = ks;
                h->fused_smem = smem;
            } else {
                cudaGetLastError();
            }
        }
    }
    if (!h->k_predict || !h->k_update || !h->k_fused) {
        set_error("no UKF kernel for model %d / sqrt %d / order %d", cfg->model_id, cfg->sqrt_mode, cfg->sigma_order);
        return fail(MPCB_BAD_ARG);
    }
    {
        int occ = 0, sms = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, reinterpret_cast<const void*>(h->k_fused), kThreads, h->fused_smem) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device) != cudaSuccess || occ < 1 || sms < 1) {
            cudaGetLastError();
            occ = 0;
        }
        h->grid_cap = (long long)occ * sms;
    }
    memset(h->Q, 0, sizeof(h->Q));
    memset(h->R, 0, sizeof(h->R));
#define TRY_OR_FAIL(expr)                                              \
    do {                                                               \
        cudaError_t _e = (expr);                                       \
        if (_e != cudaSuccess) {                                       \
            set_error("%s failed: %s", #expr, cudaGetErrorString(_e)); \
            return fail(MPCB_CUDA_ERROR);                              \
        }                                                              \
    } while (0)
    TRY_OR_FAIL(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    const size_t B = (size_t)h->B;
    TRY_OR_FAIL(cudaMalloc(&h->d_x, B * n * sizeof(double)));
    TRY_OR_FAIL(cudaMalloc(&h->d_P, B * n * n * sizeof(double)));
    TRY_OR_FAIL(cudaMalloc(&h->d_stage, (B * n * n + 128) * sizeof(double)));
    TRY_OR_FAIL(cudaMalloc(&h->d_status, B * sizeof(int)));
    TRY_OR_FAIL(cudaMemset(h->d_x, 0, B * n * sizeof(double)));
    TRY_OR_FAIL(cudaMemset(h->d_P, 0, B * n * n * sizeof(double)));
    TRY_OR_FAIL(cudaMemset(h->d_status, 0, B * sizeof(int)));
#undef TRY_OR_FAIL
    *out = h;
    return MPCB_OK;
}

mpcb_status mpcb_ukf_create(mpcb_ukf** out, const mpcb_ukf_cfg* cfg) { return ukf_create_impl(out, cfg, nullptr, nullptr, 0); }

mpcb_status mpcb_ukf_create_user(mpcb_ukf** out, const mpcb_ukf_cfg* cfg, const char* cuda_source, const double* params,
                                 int32_t n_params) {
    MPCB_REQUIRE(cuda_source != nullptr, "null source");
    MPCB_REQUIRE(n_params >= 0 && n_params <= MPCB_USER_PARAMS && (n_params == 0 || params != nullptr), "bad params");
    return ukf_create_impl(out, cfg, cuda_source, params, n_params);
}

mpcb_status mpcb_ukf_check_user_source(const char* cuda_source, int32_t n, int32_t o) {
    MPCB_REQUIRE(cuda_source != nullptr, "null source");
    MPCB_REQUIRE(n >= 1 && n <= 6 && o >= 1 && o <= 5, "user UKF models: n in 1..6, o in 1..5");
    RtcModule m;
    return rtc_compile_ukf_user(cuda_source, n, o, MPCB_SQRT_EIG, MPCB_ORDER_LIBRARY, true, false, &m);
}

void mpcb_ukf_destroy(mpcb_ukf* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    rtc_unload(&h->rtc);
    cudaFree(h->d_x);
    cudaFree(h->d_P);
    cudaFree(h->d_sigma);
    cudaFree(h->d_stage);
    cudaFree(h->d_u);
    cudaFree(h->d_z);
    cudaFree(h->d_status);
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
    delete h;
}

mpcb_status mpcb_ukf_init(mpcb_ukf* h, const double* x, const double* P, const double* Q, const double* R) {
    MPCB_REQUIRE(h && x && P && Q && R, "null pointer");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    const int n = h->n, o = h->o;
    memcpy(h->Q, Q, sizeof(double) * n * n);
    memcpy(h->R, R, sizeof(double) * o * o);
    // broadcast x and P to every filter
    MPCB_CUDA_TRY(cudaMemcpyAsync(h->d_stage, x, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    ukf_broadcast<<<blocks_for(h->B * n), 256, 0, h->stream>>>(h->d_stage, h->d_x, h->B, n);
    MPCB_CUDA_TRY(cudaGetLastError());
    MPCB_CUDA_TRY(cudaMemcpyAsync(h->d_stage + 64, P, n * n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    ukf_broadcast<<<blocks_for(h->B * n * n), 256, 0, h->stream>>>(h->d_stage + 64, h->d_P, h->B, n * n);
    MPCB_CUDA_TRY(cudaGetLastError());
    MPCB_CUDA_TRY(cudaMemsetAsync(h->d_status, 0, (size_t)h->B * sizeof(int), h->stream));
    MPCB_CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->launches += 2;
    h->predicted = false;  // sigma_f starts as NaN in the reference (src/ukf.rs:32)
    h->p_lower_only = false;
    return MPCB_OK;
}

mpcb_status mpcb_ukf_set_state(mpcb_ukf* h, const double* x, const double* P) {
    MPCB_REQUIRE(h, "null handle");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    mpcb_status st = MPCB_OK;
    if (x) {
        st = upload_aos(h, x, h->n, h->d_x);
        if (st != MPCB_OK) return st;
        MPCB_CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    if (P) {
        st = upload_aos(h, P, h->n * h->n, h->d_P);
        if (st != MPCB_OK) return st;
        h->p_lower_only = false;
    }
    MPCB_CUDA_TRY(cudaMemsetAsync(h->d_status, 0, (size_t)h->B * sizeof(int), h->stream));
    MPCB_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return MPCB_OK;
}

mpcb_status mpcb_ukf_get_state_range(mpcb_ukf* h, int64_t first, int64_t count, double* x, double* P) {
    MPCB_REQUIRE(h, "null handle");
    MPCB_REQUIRE(first >= 0 && count >= 0 && first + count <= h->B, "range outside the batch");
    if (count == 0) return MPCB_OK;
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (x) {
        ukf_soa_to_aos<<<blocks_for(count * h->n), 256, 0, h->stream>>>(h->d_x, h->d_stage, h->B, h->n, first, count);
        MPCB_CUDA_TRY(cudaGetLastError());
        MPCB_CUDA_TRY(cudaMemcpyAsync(x, h->d_stage, (size_t)count * h->n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        MPCB_CUDA_TRY(cudaStreamSynchronize(h->stream));
        h->launches += 1;
    }
    if (P) {
        const int W = h->n * h->n;
        ukf_soa_to_aos<<<blocks_for(count * W), 256, 0, h->stream>>>(h->d_P, h->d_stage, h->B, W, first, count);
        MPCB_CUDA_TRY(cudaGetLastError());
        MPCB_CUDA_TRY(cudaMemcpyAsync(P, h->d_stage, (size_t)count * W * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        MPCB_CUDA_TRY(cudaStreamSynchronize(h->stream));
        h->launches += 1;
        if (h->p_lower_only) {
            // the fused kernels store only the lower triangle of the exactly symmetric P: mirror it for the caller
            const int n = h->n;
            for (int64_t b = 0; b < count; ++b)
                for (int r = 0; r < n; ++r)
                    for (int c = r + 1; c < n; ++c) P[(b * n + r) * n + c] = P[(b * n + c) * n + r];
        }
    }
    return MPCB_OK;
}

mpcb_status mpcb_ukf_get_state(mpcb_ukf* h, double* x, double* P) {
    MPCB_REQUIRE(h, "null handle");
    return mpcb_ukf_get_state_range(h, 0, h->B, x, P);
}

mpcb_status mpcb_ukf_set_q(mpcb_ukf* h, const double* Q) {
    MPCB_REQUIRE(h && Q, "null pointer");
    memcpy(h->Q, Q, sizeof(double) * h->n * h->n);
    return MPCB_OK;
}

mpcb_status mpcb_ukf_set_r(mpcb_ukf* h, const double* R) {
    MPCB_REQUIRE(h && R, "null pointer");
    memcpy(h->R, R, sizeof(double) * h->o * h->o);
    return MPCB_OK;
}

mpcb_status mpcb_ukf_set_enable(mpcb_ukf* h, uint32_t enable) {
    MPCB_REQUIRE(h, "null handle");
    h->enable = enable;
    return MPCB_OK;
}

mpcb_status mpcb_ukf_gen_r(const mpcb_ukf* h, uint32_t enable, const double* R, double* R_out) {
    MPCB_REQUIRE(h && R && R_out, "null pointer");
    // gen_r, examples/mppi4-ukf-commu.rs:228-236: a disabled sensor's variance becomes 1e6
    for (int i = 0; i < h->o * h->o; ++i) R_out[i] = R[i];
    for (int i = 0; i < h->o; ++i)
        if ((enable & (1u << i)) == 0) R_out[i * h->o + i] = 1e6;
    return MPCB_OK;
}

mpcb_status mpcb_ukf_gather_state_device(mpcb_ukf* h, int32_t n_idx, const int32_t* idx, double* d_out) {
    MPCB_REQUIRE(h && idx && d_out, "null pointer");
    MPCB_REQUIRE(n_idx >= 1 && n_idx <= 8, "1..8 components");
    GatherParams g;
    g.x = h->d_x;
    g.out = d_out;
    g.B = h->B;
    g.n_idx = n_idx;
    for (int j = 0; j < n_idx; ++j) {
        MPCB_REQUIRE(idx[j] >= 0 && idx[j] < h->n, "component index outside the state");
        g.idx[j] = idx[j];
    }
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    ukf_gather_state_kernel<<<(unsigned)((h->B + 255) / 256), 256, 0, h->stream>>>(g);
    MPCB_CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return MPCB_OK;
}

mpcb_status mpcb_ukf_predict(mpcb_ukf* h, const double* u, double u_scalar, double dt) {
    MPCB_REQUIRE(h, "null handle");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    mpcb_status st = ensure_sigma(h);
    if (st != MPCB_OK) return st;
    UkfParams p;
    fill_params(h, &p);
    if (dt > 0.0) p.dt = dt;
    st = stage_u(h, u, u_scalar, &p);
    if (st != MPCB_OK) return st;
    st = launch(h, h->k_predict, p);
    h->p_lower_only = false;  // predict writes the whole (not exactly symmetric) P
    if (st != MPCB_OK) return st;
    MPCB_CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->predicted = true;
    return MPCB_OK;
}

mpcb_status mpcb_ukf_update(mpcb_ukf* h, const double* z) {
    MPCB_REQUIRE(h && z, "null pointer");
    if (!h->predicted) {
        set_error("update() before any predict(): sigma_f is NaN in the reference (src/ukf.rs:32)");
        return MPCB_NOT_PREDICTED;
    }
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    UkfParams p;
    fill_params(h, &p);
    mpcb_status st = stage_z(h, z, &p);
    if (st != MPCB_OK) return st;
    st = launch(h, h->k_update, p);
    if (st != MPCB_OK) return st;
    MPCB_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return MPCB_OK;
}

mpcb_status mpcb_ukf_step(mpcb_ukf* h, const double* u, double u_scalar, double dt, const double* z) {
    MPCB_REQUIRE(h && z, "null pointer");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    UkfParams p;
    fill_params(h, &p);
    if (dt > 0.0) p.dt = dt;
    mpcb_status st = stage_u(h, u, u_scalar, &p);
    if (st != MPCB_OK) return st;
    st = stage_z(h, z, &p);
    if (st != MPCB_OK) return st;
    p.lower_only = getenv("MPCB_UKF_FULL_P") ? 0 : 1;
    p.use_tma = (h->B % 4 == 0 && ((uintptr_t)p.z % 16) == 0 && (!p.has_u || ((uintptr_t)p.u % 8) == 0) && !getenv("MPCB_UKF_NO_TMA")) ? 1 : 0;
    if (p.lower_only) h->p_lower_only = true;
    st = launch(h, h->k_fused, p);
    if (st != MPCB_OK) return st;
    MPCB_CUDA_TRY(cudaStreamSynchronize(h->stream));
    // a later split update() would need sigma_f, which the fused step keeps in registers only
    h->predicted = false;
    return MPCB_OK;
}

mpcb_status mpcb_ukf_run_device(mpcb_ukf* h, int32_t steps, const double* d_u, double u_scalar, double dt,
                                const double* d_z) {
    MPCB_REQUIRE(h && d_z, "null pointer");
    MPCB_REQUIRE(steps >= 1, "steps must be >= 1");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    UkfParams p;
    fill_params(h, &p);
    if (dt > 0.0) p.dt = dt;
    p.steps = steps;
    p.u = d_u;
    p.has_u = d_u != nullptr;
    p.u_scalar = u_scalar;
    p.z = d_z;
    h->predicted = false;
    // fused step: P comes out exactly symmetric, only its lower triangle goes back to memory; the tile rows are staged with
    // TMA bulk copies when every row start is 16-byte aligned
    p.lower_only = getenv("MPCB_UKF_FULL_P") ? 0 : 1;
    p.use_tma = (h->B % 4 == 0 && ((uintptr_t)d_z % 16) == 0 && !getenv("MPCB_UKF_NO_TMA")) ? 1 : 0;
    if (p.lower_only) h->p_lower_only = true;
    return launch(h, h->k_fused, p);
}

mpcb_status mpcb_ukf_sync(mpcb_ukf* h) {
    MPCB_REQUIRE(h, "null handle");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    MPCB_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return MPCB_OK;
}

mpcb_status mpcb_ukf_get_status(mpcb_ukf* h, int32_t* s) {
    MPCB_REQUIRE(h && s, "null pointer");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    MPCB_CUDA_TRY(cudaStreamSynchronize(h->stream));
    MPCB_CUDA_TRY(cudaMemcpy(s, h->d_status, (size_t)h->B * sizeof(int), cudaMemcpyDeviceToHost));
    for (long long b = 0; b < h->B; ++b)
        if (s[b] != MPCB_OK) return (mpcb_status)s[b];
    return MPCB_OK;
}

void* mpcb_ukf_stream(mpcb_ukf* h) { return h ? (void*)h->stream : nullptr; }
int64_t mpcb_ukf_launches(mpcb_ukf* h) { return h ? h->launches : 0; }
double* mpcb_ukf_device_x(mpcb_ukf* h) { return h ? h->d_x : nullptr; }
double* mpcb_ukf_device_p(mpcb_ukf* h) { return h ? h->d_P : nullptr; }

}  // extern "C"
