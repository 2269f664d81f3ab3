// api_common.cu — status strings, error detail, model constants, defaults, device helpers.
#include <math.h>
#include <stdarg.h>

#include "common.cuh"
#include "models.cuh"

namespace mpcb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

// Folded constant prefixes, evaluated in f64 in the association order of the cited reference lines.
mpcb_status build_model_consts(int model_id, const mpcb_model_params& p, double dt, ModelConsts* out) {
    memset(out, 0, sizeof(*out));
    const double M1 = p.m1, R_W = p.r_w, M2 = p.m2, L = p.l, J1 = p.j1, J2 = p.j2, G = p.g, KT = p.kt;
    double* k = out->k;
    for (int i = 0; i < 9; ++i) k[slot::COST + i] = p.cost[i];
    switch (model_id) {
        case MPCB_MODEL_L:
        case MPCB_MODEL_PEN_LIN: {
            // examples/mppi4.rs:81-88 (same body in examples/ukf-pen.rs:14,76-83)
            const double D = (M1 + M2 + J1 / (R_W * R_W)) * (M2 * L * L + J2) - M2 * M2 * L * L;
            k[slot::L_A1] = (M1 + M2 + J1 / (R_W * R_W)) / D * M2 * G * L;
            k[slot::L_B1] = M2 * L / D / R_W * KT;
            k[slot::L_A2] = -M2 * M2 * G * L * L / D;
            k[slot::L_B2] = (M2 * L * L + J2) / D / R_W * KT;
            k[slot::L_DT] = dt;
            k[slot::L_A1DT] = k[slot::L_A1] * dt;
            k[slot::L_NB1DT] = -k[slot::L_B1] * dt;
            k[slot::L_A2DT] = k[slot::L_A2] * dt;
            k[slot::L_B2DT] = k[slot::L_B2] * dt;
            return MPCB_OK;
        }
        case MPCB_MODEL_NL:
        case MPCB_MODEL_PEN_NL:
        case MPCB_MODEL_PEN6: {
            // examples/mppi4-non-liner.rs:83-92 (same constants in ukf-pen2.rs:33-42, ukf-pen3.rs:37-48)
            k[slot::NL_D] = (M1 + M2 + J1 / (R_W * R_W)) * (M2 * L * L + J2);
            k[slot::NL_E2] = M2 * M2 * L * L;
            k[slot::NL_T1] = (M1 + M2 + J1 / (R_W * R_W)) * M2 * G * L;
            k[slot::NL_KT] = KT;
            k[slot::NL_RW] = R_W;
            k[slot::NL_ML] = M2 * L;
            k[slot::NL_M2] = M2;
            k[slot::NL_L] = L;
            k[slot::NL_JML] = J2 + M2 * L * L;
            k[slot::NL_T4] = M2 * G * L * L;
            k[slot::NL_DT] = dt;
            k[slot::NL_KTR] = KT / R_W;
            // folded constants of the FP32 form (ModelNL<float>)
            k[slot::NL_KU] = k[slot::NL_KTR] / k[slot::NL_ML];
            k[slot::NL_K3] = k[slot::NL_ML] * k[slot::NL_ML] / k[slot::NL_T1];
            k[slot::NL_K1] = k[slot::NL_JML] * k[slot::NL_ML] / k[slot::NL_T4];
            k[slot::NL_DT1] = dt * k[slot::NL_T1];
            k[slot::NL_DT4] = dt * k[slot::NL_T4];
            return MPCB_OK;
        }
        case MPCB_MODEL_NL6:
        case MPCB_MODEL_NL6_UKF: {
            // examples/mppi4-non-liner-ukf.rs:124-139
            const double B = M2 * L * L + J2;
            const double ML = M2 * L;
            const double A2 = 2.0 * M1 + M2 + 2.0 * J1 / (R_W * R_W);
            k[slot::N6_D1] = A2 * B;
            k[slot::N6_ML] = ML;
            k[slot::N6_BML] = B * M2 * L;
            k[slot::N6_NML2G] = -(ML * ML) * G;
            k[slot::N6_TWOB] = 2.0 * B;
            k[slot::N6_RW] = R_W;
            k[slot::N6_KT] = KT;
            k[slot::N6_NML2] = -(ML * ML);
            k[slot::N6_M2G] = M2 * G;
            k[slot::N6_L] = L;
            k[slot::N6_A2] = A2;
            k[slot::N6_NEG2ML] = -2.0 * M2 * L;
            k[slot::N6_DT] = dt;
            k[slot::N6_C3] = 2.0 * B * KT / R_W;
            k[slot::N6_C5] = M2 * G * L * A2;
            k[slot::N6_C6] = 2.0 * ML * KT / R_W;
            return MPCB_OK;
        }
        default:
            set_error("unknown model id %d", model_id);
            return MPCB_BAD_ARG;
    }
}

}  // namespace mpcb

using namespace mpcb;

extern "C" {

const char* mpcb_status_string(mpcb_status s) {
    switch (s) {
        case MPCB_OK: return "ok";
        case MPCB_NO_FINITE_COST: return "Cannot calculate max";
        case MPCB_SUM_ZERO: return "sum is zero";
        case MPCB_U_INVALID: return "u is invalid";
        case MPCB_INVERSE_FAIL: return "Inverse fail";
        case MPCB_CHOLESKY_FAIL: return "Cholesky fail";
        case MPCB_BAD_ARG: return "bad argument";
        case MPCB_CUDA_ERROR: return "CUDA error";
        case MPCB_NCCL_ERROR: return "NCCL error";
        case MPCB_NOT_PREDICTED: return "update before predict";
        case MPCB_PEER_TIMEOUT: return "peer exchange timed out";
        case MPCB_RTC_ERROR: return "user model did not compile";
        default: return "unknown status";
    }
}

const char* mpcb_last_error_string(void) { return last_error(); }
int mpcb_abi_version(void) { return MPCB_ABI_VERSION; }

int mpcb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

mpcb_status mpcb_model_defaults(int32_t model_id, mpcb_model_params* p) {
    if (!p) return MPCB_BAD_ARG;
    memset(p, 0, sizeof(*p));
    switch (model_id) {
        case MPCB_MODEL_L:       // examples/mppi4.rs:8-18,73-81
        case MPCB_MODEL_NL:      // examples/mppi4-non-liner.rs:8-18,73-80
        case MPCB_MODEL_PEN_LIN: // examples/ukf-pen.rs:6-16
        case MPCB_MODEL_PEN_NL:  // examples/ukf-pen2.rs:8-17
        case MPCB_MODEL_PEN6:    // examples/ukf-pen3.rs:8-17
            p->m1 = 150e-3;
            p->r_w = 50e-3;
            p->m2 = 2.3 - 2.0 * p->m1 + 2.0;
            p->l = 0.2474;
            p->j1 = p->m1 * p->r_w * p->r_w;
            p->j2 = (model_id == MPCB_MODEL_PEN_LIN) ? 0.1 : 0.2;
            p->g = 9.81;
            p->kt = 0.15;
            if (model_id == MPCB_MODEL_L || model_id == MPCB_MODEL_NL) {
                p->dt = 0.8 / 8.0;
                const double cw[9] = {2.0, 3.0, 5.0, 1.2, 2.0, 5.0, 2.0, 0.35, 0.75};
                for (int i = 0; i < 9; ++i) p->cost[i] = cw[i];
            } else {
                p->dt = 0.01;
            }
            return MPCB_OK;
        case MPCB_MODEL_NL6:     // examples/mppi4-non-liner-ukf.rs:13-24,108-124
        case MPCB_MODEL_NL6_UKF:
            p->m1 = 160e-3;
            p->r_w = 50e-3;
            p->m2 = 2.4;
            p->l = 0.4;
            p->j1 = 2.23e5 * 1e-9;
            p->j2 = 1.168e8 * 1e-9;
            p->g = 9.81;
            p->kt = 0.15;
            p->dt = 1.2 / 8.0;
            p->cost[0] = 0.1; p->cost[1] = 0.1; p->cost[2] = 1.0; p->cost[3] = 0.5;
            return MPCB_OK;
        default:
            set_error("unknown model id %d", model_id);
            return MPCB_BAD_ARG;
    }
}

mpcb_status mpcb_device_alloc(int32_t device, uint64_t bytes, void** out) {
    MPCB_REQUIRE(out != nullptr, "out is null");
    MPCB_CUDA_TRY(cudaSetDevice(device));
    MPCB_CUDA_TRY(cudaMalloc(out, bytes ? bytes : 1));
    return MPCB_OK;
}
mpcb_status mpcb_device_free(int32_t device, void* p) {
    MPCB_CUDA_TRY(cudaSetDevice(device));
    MPCB_CUDA_TRY(cudaFree(p));
    return MPCB_OK;
}
mpcb_status mpcb_device_upload(int32_t device, void* dst, const void* src, uint64_t bytes) {
    MPCB_CUDA_TRY(cudaSetDevice(device));
    MPCB_CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return MPCB_OK;
}
mpcb_status mpcb_device_download(int32_t device, void* dst, const void* src, uint64_t bytes) {
    MPCB_CUDA_TRY(cudaSetDevice(device));
    MPCB_CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return MPCB_OK;
}

}  // extern "C"
