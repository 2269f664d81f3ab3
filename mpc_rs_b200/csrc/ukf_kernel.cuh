// ukf_kernel.cuh — batched unscented Kalman filter, one thread per filter, everything in registers, FP64.
//
// Replaces UnscentedKalmanFilter::{predict,update} of src/ukf.rs:44-74 (n=4,o=3), src/ukf2.rs:44-74
// (n=6,o=5) and the free functions of examples/ukf-pen.rs:44-141 (n=4,o=2, Cholesky).  The unscented
// weights are +-1e6 (alpha = 1e-3, src/ukf.rs:24-28), which amplifies rounding by ~1.7e5 per step: FP32 is
// unusable here (SURVEY.md finding 4), so state, covariance, sigma points and the model functions are FP64
// and the arithmetic follows the reference's association order (this TU is compiled with -fmad=false).
//
// HBM layout is structure-of-arrays [component][B]: every load/store of a warp is one 256-byte coalesced
// transaction per component.  A fused step reads x, P, z and writes x, P: 8*(2n + 2n^2 + o) bytes per
// filter-update (336 B at n=4,o=2); with steps > 1 the state stays in registers between steps and only z
// streams in.
#pragma once

#include <math_constants.h>

#include "common.cuh"
#include "models.cuh"

namespace mpcb {

namespace uslot {  // hx constants, disjoint from the dynamics slots of models.cuh
constexpr int G = 12, L = 13, RPM = 14, NRPM = 15, DEG = 17, M2G = 18, M2 = 19, M2L = 20;
constexpr int IG = 29;  // 1 / g (fast arithmetic: the accelerometer rows multiply instead of dividing)
}

struct UkfParams {
    long long B;
    int steps;
    int has_u;
    double* x;        // [n][B]
    double* P;        // [n*n][B]
    double* sigma_f;  // [n*M][B] (split predict/update only)
    const double* u;  // [steps][B] or nullptr
    const double* z;  // [steps][o][B]
    int* status;      // [B], sticky
    double u_scalar;
    double dt;
    unsigned int enable;  // sensor bit mask: hx rows of cleared bits read 0 (examples/mppi4-ukf-commu.rs:279-293)
    int use_tma;           // fused n = 4 kernels: stage the tile rows with cp.async.bulk (needs B % 4 == 0 and 16-byte aligned x, P, z, status)
    int lower_only;        // fused kernels: store only the lower triangle of the (exactly symmetric) P; readers mirror it
    unsigned int reverse;  // 1: walk the tiles from the end (alternates per launch: the tiles the previous launch wrote last are still in L2)
    double wm0, wc0, wi, cC;  // sigma_weight (src/ukf.rs:112-118), C = alpha^2 (n + kappa)
    double Q[36];
    double R[25];
    ModelConsts mc;
};

enum UkfMode { UKF_PREDICT = 0, UKF_UPDATE = 1, UKF_FUSED = 2 };

// ------------------------------------------------------------------------------------------------
// process / measurement models (f64, reference association order)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void nl6_ddot(const ModelConsts& mc, double th, double thd, double u, double* ddx,
                                         double* ddth) {
    // examples/mppi4-non-liner-ukf.rs:126-139 with f = 0 (the UKF process model passes f = 0, :278)
    const double s2 = sin(th), c2 = cos(th);
    const double mlc = mc.k[slot::N6_ML] * c2;
    const double d = mc.k[slot::N6_D1] - mlc * mlc;
    const double w2 = thd * thd;
    const double term1 = mc.k[slot::N6_BML] / d * w2 * s2;
    const double term2 = mc.k[slot::N6_NML2G] / d * s2 * c2;
    const double term3 = mc.k[slot::N6_TWOB] / (d * mc.k[slot::N6_RW]) * mc.k[slot::N6_KT] * u;
    *ddx = term1 + term2 + term3 + 0.0;
    const double t1 = mc.k[slot::N6_NML2] / d * w2 * s2 * c2;
    const double t2 = (mc.k[slot::N6_M2G] * s2 - 2.0 * 0.0) * mc.k[slot::N6_L] * mc.k[slot::N6_A2] / d;
    const double t3 = mc.k[slot::N6_NEG2ML] / (d * mc.k[slot::N6_RW]) * mc.k[slot::N6_KT] * u * c2;
    *ddth = t1 + t2 + t3 + 0.0;
}

// The same accelerations in the folded form of ModelNL6F (models.cuh): one sincos, ONE reciprocal instead of five IEEE
// divisions, FMA — the fast-arithmetic kernels (rounding-level differences from the reference order)
__device__ __forceinline__ void nl6_ddot_fast(const ModelConsts& mc, double th, double thd, double u, double* ddx, double* ddth) {
    double s2, c2;
    sincos_f64_fast(th, &s2, &c2);
    const double mlc = mc.k[slot::N6_ML] * c2;
    const double id = rcp_f64_fast(fma(-mlc, mlc, mc.k[slot::N6_D1]));
    const double ws = (thd * thd) * s2;
    const double numx = fma(mc.k[slot::N6_BML], ws, fma(mc.k[slot::N6_NML2G] * s2, c2, mc.k[slot::N6_C3] * u));
    const double numt = fma(c2, fma(mc.k[slot::N6_NML2], ws, -mc.k[slot::N6_C6] * u), mc.k[slot::N6_C5] * s2);
    *ddx = numx * id;
    *ddth = numt * id;
}

template <int MODEL, int N, bool FAST = false>
__device__ __forceinline__ void ukf_fx(const ModelConsts& mc, double (&x)[N], double u, double dt) {
#ifdef MPCB_UKF_USER
    // run-time compiled user model (mpcb_ukf_create_user): ::fx is declared by the user source ahead of this header
    if constexpr (MODEL == MPCB_MODEL_USER_UKF) {
        ::fx(x, u, dt, mc.k);
    } else
#endif
    if constexpr (MODEL == MPCB_MODEL_PEN_LIN) {
        // examples/ukf-pen.rs:76-83
        x[3] += (mc.k[slot::L_A1] * x[2] - mc.k[slot::L_B1] * u) * dt;
        x[2] += x[3] * dt;
        x[1] += (mc.k[slot::L_A2] * x[2] + mc.k[slot::L_B2] * u) * dt;
        x[0] += x[1] * dt;
    } else if constexpr (MODEL == MPCB_MODEL_PEN_NL && FAST) {
        // the folded form of ModelNLF (models.cuh): one sincos, one reciprocal, FMA
        double s, c;
        sincos_f64_fast(x[2], &s, &c);
        const double rd = rcp_f64_fast(fma(-mc.k[slot::NL_E2], c * c, mc.k[slot::NL_D]));
        const double qq = fma(mc.k[slot::NL_KU], u, (x[3] * x[3]) * s);
        const double n3 = fma(-mc.k[slot::NL_K3], qq * c, s);
        const double n1 = fma(mc.k[slot::NL_K1], qq, s * c);
        const double r3 = fma(n3, rd * (dt * mc.k[slot::NL_T1]), x[3]);
        const double r2 = fma(x[3], dt, x[2]);
        const double r1 = fma(n1, rd * (dt * mc.k[slot::NL_T4]), x[1]);
        const double r0 = fma(x[1], dt, x[0]);
        x[0] = r0; x[1] = r1; x[2] = r2; x[3] = r3;
    } else if constexpr (MODEL == MPCB_MODEL_PEN_NL) {
        // examples/ukf-pen2.rs:31-44
        const double s = sin(x[2]), c = cos(x[2]);
        const double d = mc.k[slot::NL_D] - mc.k[slot::NL_E2] * c * c;
        const double term1 = mc.k[slot::NL_T1] * s;
        const double q = mc.k[slot::NL_KT] * u / mc.k[slot::NL_RW] + mc.k[slot::NL_ML] * (x[3] * x[3]) * s;
        const double term2 = q * mc.k[slot::NL_M2] * mc.k[slot::NL_L] * c;
        const double r3 = x[3] + (term1 - term2) / d * dt;
        const double r2 = x[2] + x[3] * dt;
        const double term3 = mc.k[slot::NL_JML] * q;
        const double term4 = mc.k[slot::NL_T4] * s * c;
        const double r1 = x[1] + (term3 + term4) / d * dt;
        const double r0 = x[0] + x[1] * dt;
        x[0] = r0; x[1] = r1; x[2] = r2; x[3] = r3;
    } else if constexpr (MODEL == MPCB_MODEL_PEN6 && FAST) {
        // same formulas, one sincos and one reciprocal (d still comes from x[2].cos(), as the reference writes it)
        const double mlc = mc.k[slot::NL_ML] * cos(x[2]);
        const double id = rcp_f64_fast(fma(-mlc, mlc, mc.k[slot::NL_D]));
        double s3, c3;
        sincos_f64_fast(x[3], &s3, &c3);
        const double q = fma(mc.k[slot::NL_ML] * (x[4] * x[4]), s3, mc.k[slot::NL_KTR] * u);
        const double r0 = fma(x[1], dt, x[0]);
        const double r1 = fma(x[2], dt, x[1]);
        const double r2 = fma(mc.k[slot::NL_JML], q, mc.k[slot::NL_T4] * s3 * c3) * id;
        const double r3 = fma(x[4], dt, x[3]);
        const double r4 = fma(x[5], dt, x[4]);
        const double r5 = fma(-(mc.k[slot::NL_M2] * mc.k[slot::NL_L]) * q, c3, mc.k[slot::NL_T1] * s3) * id;
        x[0] = r0; x[1] = r1; x[2] = r2; x[3] = r3; x[4] = r4; x[5] = r5;
    } else if constexpr (MODEL == MPCB_MODEL_PEN6) {
        // examples/ukf-pen3.rs:35-50 — d from x[2].cos() as written (theta is x[3] in this layout)
        const double mlc = mc.k[slot::NL_ML] * cos(x[2]);
        const double d = mc.k[slot::NL_D] - mlc * mlc;
        const double s3 = sin(x[3]), c3 = cos(x[3]);
        const double q = mc.k[slot::NL_KT] * u / mc.k[slot::NL_RW] + mc.k[slot::NL_ML] * (x[4] * x[4]) * s3;
        const double r0 = x[0] + x[1] * dt;
        const double r1 = x[1] + x[2] * dt;
        const double term3 = mc.k[slot::NL_JML] * q;
        const double term4 = mc.k[slot::NL_T4] * s3 * c3;
        const double r2 = (term3 + term4) / d;
        const double r3 = x[3] + x[4] * dt;
        const double r4 = x[4] + x[5] * dt;
        const double term1 = mc.k[slot::NL_T1] * s3;
        const double term2 = q * mc.k[slot::NL_M2] * mc.k[slot::NL_L] * c3;
        const double r5 = (term1 - term2) / d;
        x[0] = r0; x[1] = r1; x[2] = r2; x[3] = r3; x[4] = r4; x[5] = r5;
    } else {
        // MPCB_MODEL_NL6_UKF: dynamics_short(x, u, dt, 0) — examples/mppi4-non-liner-ukf.rs:149-159
        double ddx, ddth;
        if constexpr (FAST) nl6_ddot_fast(mc, x[3], x[4], u, &ddx, &ddth);
        else nl6_ddot(mc, x[3], x[4], u, &ddx, &ddth);
        x[5] = ddth;
        x[4] += x[5] * dt;
        x[3] += x[4] * dt;
        x[2] = ddx;
        x[1] += x[2] * dt;
        x[0] += x[1] * dt;
    }
}

template <int MODEL, int N, int O, bool FAST = false>
__device__ __forceinline__ void ukf_hx(const ModelConsts& mc, const double (&x)[N], double (&z)[O]) {
#ifdef MPCB_UKF_USER
    if constexpr (MODEL == MPCB_MODEL_USER_UKF) {
        ::hx(x, z, mc.k);
    } else
#endif
    if constexpr (MODEL == MPCB_MODEL_PEN_LIN) {
        z[0] = x[1];  // examples/ukf-pen.rs:86-91
        z[1] = x[3];
    } else if constexpr (MODEL == MPCB_MODEL_PEN_NL) {
        z[0] = mc.k[uslot::RPM] * x[1];  // examples/ukf-pen2.rs:47-53
        z[1] = mc.k[uslot::RPM] * x[1];
        z[2] = x[3] * mc.k[uslot::DEG];
    } else if constexpr (MODEL == MPCB_MODEL_PEN6 && FAST) {
        double s3, c3;
        sincos_f64_fast(x[3], &s3, &c3);
        const double m2a = mc.k[uslot::M2] * x[2];
        const double v = fma(mc.k[uslot::M2G], c3, fma(m2a, s3, -mc.k[uslot::M2L] * (x[4] * x[4])));
        const double h = fma(-mc.k[uslot::M2G], s3, fma(m2a, c3, mc.k[uslot::M2L] * x[5]));
        z[0] = mc.k[uslot::RPM] * x[1];
        z[1] = z[0];
        z[2] = x[3] * mc.k[uslot::DEG];
        z[3] = v * mc.k[uslot::IG];
        z[4] = h * mc.k[uslot::IG];
    } else if constexpr (MODEL == MPCB_MODEL_PEN6) {
        // examples/ukf-pen3.rs:53-63
        const double s3 = sin(x[3]), c3 = cos(x[3]);
        const double v = mc.k[uslot::M2G] * c3 + mc.k[uslot::M2] * x[2] * s3 - mc.k[uslot::M2L] * (x[4] * x[4]);
        const double h = -mc.k[uslot::M2G] * s3 + mc.k[uslot::M2] * x[2] * c3 + mc.k[uslot::M2L] * x[5];
        z[0] = mc.k[uslot::RPM] * x[1];
        z[1] = mc.k[uslot::RPM] * x[1];
        z[2] = x[3] * mc.k[uslot::DEG];
        z[3] = v / mc.k[uslot::G];
        z[4] = h / mc.k[uslot::G];
    } else if constexpr (FAST) {
        double s3, c3;
        sincos_f64_fast(x[3], &s3, &c3);
        const double ax = fma(mc.k[uslot::G], s3, fma(x[2], c3, mc.k[uslot::L] * x[5]));
        const double az = fma(mc.k[uslot::G], c3, fma(-x[2], s3, mc.k[uslot::L] * (x[4] * x[4])));
        z[0] = mc.k[uslot::RPM] * x[1];
        z[1] = mc.k[uslot::NRPM] * x[1];
        z[2] = x[4] * mc.k[uslot::DEG];
        z[3] = az * mc.k[uslot::IG];
        z[4] = ax * mc.k[uslot::IG];
    } else {
        // examples/mppi4-non-liner-ukf.rs:169-179
        const double s3 = sin(x[3]), c3 = cos(x[3]);
        const double ax = mc.k[uslot::G] * s3 + x[2] * c3 + mc.k[uslot::L] * x[5];
        const double az = mc.k[uslot::G] * c3 - x[2] * s3 + mc.k[uslot::L] * (x[4] * x[4]);
        z[0] = mc.k[uslot::RPM] * x[1];
        z[1] = mc.k[uslot::NRPM] * x[1];
        z[2] = x[4] * mc.k[uslot::DEG];
        z[3] = az / mc.k[uslot::G];
        z[4] = ax / mc.k[uslot::G];
    }
}

// Out-of-line calls for the six-state models: their fx/hx carry FP64 sin/cos expansions, and inlining them at all
// 2n+1 = 13 sigma points made the fused kernel 220 KB of straight-line code — ncu: "no_instruction" was the top
// stall (2.8 cycles per issue), the FP64 pipe 23 % busy.  The four-state kernels stay inlined (their code fits).
template <int MODEL, int N, bool FAST>
__device__ __noinline__ void ukf_fx_call(const ModelConsts& mc, double (&x)[N], double u, double dt) {
    ukf_fx<MODEL, N, FAST>(mc, x, u, dt);
}
template <int MODEL, int N, int O, bool FAST>
__device__ __noinline__ void ukf_hx_call(const ModelConsts& mc, const double (&x)[N], double (&z)[O]) {
    ukf_hx<MODEL, N, O, FAST>(mc, x, z);
}

// ------------------------------------------------------------------------------------------------
// small dense kernels, fully unrolled
// ------------------------------------------------------------------------------------------------
// nalgebra-style lower Cholesky (reads the lower triangle); false if a pivot is not > 0
// FAST: one rsqrt per pivot instead of a sqrt and N-j-1 divisions on the serial chain (rounding-level differences)
template <int N, bool FAST = false>
__device__ __forceinline__ bool chol_lower(double (&m)[N][N]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < N; ++j) {
#pragma unroll
        for (int k = 0; k < j; ++k) {
            const double ljk = m[j][k];
#pragma unroll
            for (int i = j; i < N; ++i) m[i][j] -= m[i][k] * ljk;
        }
        const double diag = m[j][j];
        if (!(diag > 0.0)) ok = false;
        if constexpr (FAST) {
            const double inv = rsqrt(diag);
            m[j][j] = diag * inv;
#pragma unroll
            for (int i = j + 1; i < N; ++i) m[i][j] *= inv;
        } else {
            const double denom = sqrt(diag);
            m[j][j] = denom;
#pragma unroll
            for (int i = j + 1; i < N; ++i) m[i][j] /= denom;
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i + 1; j < N; ++j) m[i][j] = 0.0;
    return ok;
}

// U*sqrt(S) of the SVD of a symmetric PSD matrix by cyclic Jacobi (src/ukf.rs:121-124); at most 10
// row-cyclic sweeps, stops when every off-diagonal element is exactly zero.
template <int N, bool FAST = false>
__device__ __forceinline__ void sym_eig_sqrt(double (&A)[N][N], double (&Lo)[N][N]) {
    double V[N][N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            V[i][j] = (i == j) ? 1.0 : 0.0;
            if (j > i) A[i][j] = A[j][i];  // lower triangle is the input
        }
    for (int sweep = 0; sweep < 10; ++sweep) {
        double off = 0.0;
        double skip = 0.0;  // FAST: rotations whose a_pq^2 is below their share of the stopping mass are left out
#pragma unroll
        for (int p = 0; p < N - 1; ++p)
#pragma unroll
            for (int q = p + 1; q < N; ++q) off += A[p][q] * A[p][q];
        if constexpr (FAST) {
            // converged when the off-diagonal mass is below rounding of the diagonal: sum a_pq^2 <= (eps/4)^2 sum a_kk^2.
            // Jacobi converges quadratically, so the exact-zero criterion of the reference-order variant costs three to
            // four further sweeps (1e-32 -> 1e-64 -> ... -> underflow) that change nothing above 1e-16 relative.
            double dsum = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) dsum = fma(A[k][k], A[k][k], dsum);
            if (off <= 3.0e-33 * dsum) break;
            // an entry with a_pq^2 <= (3e-33 / #pairs) sum a_kk^2 cannot keep the sweep loop alive on its own: if every
            // entry is that small the criterion above holds.  Leaving those rotations out changes nothing above rounding
            // and, in the last sweep or two, most rotations of most filters of a warp are that small.
            skip = (3.0e-33 / (N * (N - 1) / 2)) * dsum;
        } else {
            if (off == 0.0) break;
        }
#pragma unroll
        for (int p = 0; p < N - 1; ++p) {
#pragma unroll
            for (int q = p + 1; q < N; ++q) {
                const double apq = A[p][q];
                bool rotate;
                if constexpr (FAST) rotate = apq * apq > skip;
                else rotate = apq != 0.0;
                if (rotate) {
                    if constexpr (!FAST) {
                        // the oracle's operation order (two divisions, two square roots, full two-sided update)
                        const double tau = (A[q][q] - A[p][p]) / (2.0 * apq);
                        const double rt = sqrt(1.0 + tau * tau);
                        const double t = (tau >= 0.0) ? 1.0 / (tau + rt) : -1.0 / (-tau + rt);
                        const double c = 1.0 / sqrt(1.0 + t * t);
                        const double s = t * c;
#pragma unroll
                        for (int k = 0; k < N; ++k) {
                            const double akp = A[k][p], akq = A[k][q];
                            A[k][p] = c * akp - s * akq;
                            A[k][q] = s * akp + c * akq;
                        }
#pragma unroll
                        for (int k = 0; k < N; ++k) {
                            const double apk = A[p][k], aqk = A[q][k];
                            A[p][k] = c * apk - s * aqk;
                            A[q][k] = s * apk + c * aqk;
                        }
                        A[p][q] = 0.0;
                        A[q][p] = 0.0;
#pragma unroll
                        for (int k = 0; k < N; ++k) {
                            const double vkp = V[k][p], vkq = V[k][q];
                            V[k][p] = c * vkp - s * vkq;
                            V[k][q] = s * vkp + c * vkq;
                        }
                    } else {
                        // same rotation, symmetric update (only rows/columns p and q of the other indices change, the 2x2 block
                        // in closed form: a_pp -= t a_pq, a_qq += t a_pq, a_pq = 0), and the angle from TWO reciprocal square
                        // roots and no division: with theta = (a_qq - a_pp)/2 and h = hypot(theta, a_pq),
                        //   cos 2phi = |theta|/h,  cos^2 phi = (1 + cos 2phi)/2,  sin phi = (|a_pq|/h) / (2 cos phi),  t = sin/cos
                        // — the same phi as t = sign(tau)/(|tau| + sqrt(1 + tau^2)), tau = theta/a_pq.  Equal diagonal entries
                        // rotate by +45 degrees whatever the sign of a_pq, like the division form does.
                        const double th = 0.5 * (A[q][q] - A[p][p]);
                        const double h2 = fma(th, th, apq * apq);
                        double c, s, t;
                        if (h2 > 1e-280 && h2 < 1e280) {
                            const double rh = rsqrt_f64_fast(h2);
                            const double c2 = fma(0.5, fabs(th) * rh, 0.5);  // in [0.5, 1]
                            const double rc = rsqrt_f64_fast(c2);
                            c = c2 * rc;
                            const double sa = (0.5 * rc) * (fabs(apq) * rh);
                            s = (th != 0.0 && ((th < 0.0) != (apq < 0.0))) ? -sa : sa;
                            t = s * rc;
                        } else {
                            // h^2 under- or overflows (entries around 1e-140 or 1e140): the division form
                            const double tau = (th == 0.0) ? 0.0 : th * __drcp_rn(apq);
                            const double at = fabs(tau);
                            if (at < 1e150) {
                                const double w = fma(tau, tau, 1.0);
                                t = copysign(__drcp_rn(at + w * rsqrt(w)), tau);
                            } else {
                                t = 0.5 * __drcp_rn(tau);
                            }
                            c = rsqrt(fma(t, t, 1.0));
                            s = t * c;
                        }
#pragma unroll
                        for (int k = 0; k < N; ++k) {
                            if (k != p && k != q) {
                                const double akp = A[k][p], akq = A[k][q];
                                const double np_ = c * akp - s * akq, nq_ = s * akp + c * akq;
                                A[k][p] = np_; A[p][k] = np_;
                                A[k][q] = nq_; A[q][k] = nq_;
                            }
                        }
                        A[p][p] -= t * apq;
                        A[q][q] += t * apq;
                        A[p][q] = 0.0;
                        A[q][p] = 0.0;
#pragma unroll
                        for (int k = 0; k < N; ++k) {
                            const double vkp = V[k][p], vkq = V[k][q];
                            V[k][p] = c * vkp - s * vkq;
                            V[k][q] = s * vkp + c * vkq;
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const double sq = sqrt(fabs(A[j][j]));
#pragma unroll
        for (int i = 0; i < N; ++i) Lo[i][j] = V[i][j] * sq;
    }
}

// try_inverse: closed forms for 2x2 / 3x3, LU with partial pivoting otherwise; false = singular
template <int O, bool FAST = false>
__device__ __forceinline__ bool inverse_small(const double (&A)[O][O], double (&Ai)[O][O]) {
    if constexpr (O == 2) {
        const double det = A[0][0] * A[1][1] - A[1][0] * A[0][1];
        if (det == 0.0) return false;
        if constexpr (FAST) {  // one reciprocal instead of four divisions
            const double id = __drcp_rn(det);
            Ai[0][0] = A[1][1] * id; Ai[0][1] = -A[0][1] * id; Ai[1][0] = -A[1][0] * id; Ai[1][1] = A[0][0] * id;
        } else {
            Ai[0][0] = A[1][1] / det; Ai[0][1] = -A[0][1] / det; Ai[1][0] = -A[1][0] / det; Ai[1][1] = A[0][0] / det;
        }
        return true;
    } else if constexpr (O == 3) {
        const double m11 = A[0][0], m12 = A[0][1], m13 = A[0][2], m21 = A[1][0], m22 = A[1][1], m23 = A[1][2],
                     m31 = A[2][0], m32 = A[2][1], m33 = A[2][2];
        const double minor_m12_m23 = m22 * m33 - m32 * m23;
        const double minor_m11_m23 = m21 * m33 - m31 * m23;
        const double minor_m11_m22 = m21 * m32 - m31 * m22;
        const double det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
        if (det == 0.0) return false;
        Ai[0][0] = minor_m12_m23 / det;
        Ai[0][1] = (m13 * m32 - m33 * m12) / det;
        Ai[0][2] = (m12 * m23 - m22 * m13) / det;
        Ai[1][0] = -minor_m11_m23 / det;
        Ai[1][1] = (m11 * m33 - m31 * m13) / det;
        Ai[1][2] = (m13 * m21 - m23 * m11) / det;
        Ai[2][0] = minor_m11_m22 / det;
        Ai[2][1] = (m12 * m31 - m32 * m11) / det;
        Ai[2][2] = (m11 * m22 - m21 * m12) / det;
        return true;
    } else {
        double lu[O][O], rhs[O][O];  // rhs = row-permuted identity
        bool ok = true;
#pragma unroll
        for (int i = 0; i < O; ++i)
#pragma unroll
            for (int j = 0; j < O; ++j) {
                lu[i][j] = A[i][j];
                rhs[i][j] = (i == j) ? 1.0 : 0.0;
            }
        // (full-range loops with compile-time guards: nvcc left the triangular forms partly rolled, with lu / rhs in local memory)
#pragma unroll
        for (int k = 0; k < O; ++k) {
            int piv = k;
            double best = fabs(lu[k][k]);
#pragma unroll
            for (int i = 0; i < O; ++i) {
                if (i > k) {
                    const double v = fabs(lu[i][k]);
                    if (v > best) { best = v; piv = i; }
                }
            }
            if (best == 0.0) ok = false;
#pragma unroll
            for (int i = 0; i < O; ++i) {
                if (i > k) {
                    const bool sw = (piv == i);
#pragma unroll
                    for (int j = 0; j < O; ++j) {
                        const double a = lu[k][j], b = lu[i][j];
                        lu[k][j] = sw ? b : a;
                        lu[i][j] = sw ? a : b;
                        const double ra = rhs[k][j], rb = rhs[i][j];
                        rhs[k][j] = sw ? rb : ra;
                        rhs[i][j] = sw ? ra : rb;
                    }
                }
            }
            const double d = lu[k][k];
#pragma unroll
            for (int i = 0; i < O; ++i) {
                if (i > k) {
                    lu[i][k] /= d;
                    const double l = lu[i][k];
#pragma unroll
                    for (int j = 0; j < O; ++j)
                        if (j > k) lu[i][j] -= l * lu[k][j];
                }
            }
        }
#pragma unroll
        for (int c = 0; c < O; ++c) {
            double y[O];
#pragma unroll
            for (int i = 0; i < O; ++i) {
                double s = rhs[i][c];
#pragma unroll
                for (int j = 0; j < O; ++j)
                    if (j < i) s -= lu[i][j] * y[j];
                y[i] = s;
            }
#pragma unroll
            for (int ii = 0; ii < O; ++ii) {
                constexpr int kLast = O - 1;
                const int i = kLast - ii;
                double s = y[i];
#pragma unroll
                for (int j = 0; j < O; ++j)
                    if (j > i) s -= lu[i][j] * Ai[j][c];
                Ai[i][c] = s / lu[i][i];
            }
        }
        return ok;
    }
}

// sigma-point column index of +L_i / -L_i
template <int N, int ORDER>
__device__ __forceinline__ constexpr int col_plus(int i) { return ORDER == MPCB_ORDER_INTERLEAVED ? 1 + 2 * i : 1 + i; }
template <int N, int ORDER>
__device__ __forceinline__ constexpr int col_minus(int i) { return ORDER == MPCB_ORDER_INTERLEAVED ? 2 + 2 * i : 1 + N + i; }

// unscented_transform (src/ukf.rs:96-110): mean = sigmas*wm, P = sum_i wc_i y_i y_i^T + cov
//   FAST = false: the reference's operation order, every (r,c) entry accumulated separately, no FMA (TU flag)
//   FAST = true : the weight shared by i >= 1 is factored out, only the upper triangle is accumulated and then
//                 mirrored (each y_i y_i^T is symmetric), FMA contraction on — ~40% fewer FP64 instructions;
//                 differs from the reference order at rounding level (x 1.7e5 weight amplification ~ 1e-10)
template <int S, int M, bool FAST>
__device__ __forceinline__ void unscented_transform(const double (&sig)[S][M], double wm0, double wc0, double wi,
                                                    const double* cov, double (&mean)[S], double (&P)[S][S]) {
    if constexpr (!FAST) {
#pragma unroll
        for (int r = 0; r < S; ++r) {
            double acc = sig[r][0] * wm0;
#pragma unroll
            for (int i = 1; i < M; ++i) acc += sig[r][i] * wi;
            mean[r] = acc;
        }
#pragma unroll
        for (int r = 0; r < S; ++r)
#pragma unroll
            for (int c = 0; c < S; ++c) P[r][c] = 0.0;
#pragma unroll
        for (int i = 0; i < M; ++i) {
            double y[S];
#pragma unroll
            for (int r = 0; r < S; ++r) y[r] = sig[r][i] - mean[r];
            const double w = (i == 0) ? wc0 : wi;
#pragma unroll
            for (int r = 0; r < S; ++r) {
                const double wy = w * y[r];
#pragma unroll
                for (int c = 0; c < S; ++c) P[r][c] += wy * y[c];
            }
        }
#pragma unroll
        for (int r = 0; r < S; ++r)
#pragma unroll
            for (int c = 0; c < S; ++c) P[r][c] = P[r][c] + cov[r * S + c];
    } else {
#pragma unroll
        for (int r = 0; r < S; ++r) {
            double acc = sig[r][1];
#pragma unroll
            for (int i = 2; i < M; ++i) acc += sig[r][i];
            mean[r] = fma(acc, wi, sig[r][0] * wm0);
        }
#pragma unroll
        for (int r = 0; r < S; ++r)
#pragma unroll
            for (int c = r; c < S; ++c) P[r][c] = 0.0;
#pragma unroll
        for (int i = 1; i < M; ++i) {
            double y[S];
#pragma unroll
            for (int r = 0; r < S; ++r) y[r] = sig[r][i] - mean[r];
#pragma unroll
            for (int r = 0; r < S; ++r)
#pragma unroll
                for (int c = r; c < S; ++c) P[r][c] = fma(y[r], y[c], P[r][c]);
        }
        double y0[S];
#pragma unroll
        for (int r = 0; r < S; ++r) y0[r] = sig[r][0] - mean[r];
#pragma unroll
        for (int r = 0; r < S; ++r) {
            const double wy = wc0 * y0[r];
#pragma unroll
            for (int c = r; c < S; ++c) {
                const double v = fma(wi, P[r][c], fma(wy, y0[c], cov[r * S + c]));
                P[r][c] = v;
                P[c][r] = v;
            }
        }
    }
}

// 8- and 4-byte asynchronous global -> shared copies (LDGSTS): a thread stages ITS OWN filter's inputs of the next tile
// while it computes the current one, so no barrier is needed — only the thread's own cp.async.wait_group.
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int KEEP>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(KEEP) : "memory"); }

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier: ONE thread stages a whole SoA row of the tile ----
__device__ __forceinline__ void ukf_mbar_init(unsigned long long* bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ukf_mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ukf_mbar_wait(unsigned long long* bar, unsigned int parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "UW_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra UD_%=;\n"
        "bra UW_%=;\n"
        "UD_%=:\n"
        "}\n" ::"r"(addr), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void ukf_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned int bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

constexpr int kUkfThreads = 128;

// One thread per filter, a block walks tiles of 128 filters (grid-stride).  The fused predict+update kernel of the
// four-state filters is bound by HBM *latency* when every block loads, computes and stores in sequence (ncu: 12 warps
// per SM at 158 registers, top stall long_scoreboard, DRAM 46 % and FP64 pipe 47 % busy): there the inputs of tile
// t+1 (x, P, z of step 0, status: 23 rows of 128 values) are copied into a second shared-memory buffer with cp.async
// while tile t is computed (PIPE).  The other modes and n = 6 (FP64-bound, 2x the rows) load directly.
template <int N, int O, int MODEL, int SQRT, int ORDER, int MODE, bool FAST>
__global__ void __launch_bounds__(kUkfThreads) ukf_kernel(const __grid_constant__ UkfParams p) {
    pdl_entry();
    constexpr int M = 2 * N + 1;
    // predict never reads the strictly-upper triangle of P (the Cholesky / eigen square root take the lower one, the
    // unscented transform then rewrites P entirely), so the fused and predict kernels do not load it: 288 instead of
    // 336 bytes of DRAM traffic per n = 4, o = 2 update, same result bit for bit
    constexpr int NTRI = N * (N + 1) / 2;
    constexpr int ROWS = N + NTRI + O;
    constexpr bool PIPE = (MODE == UKF_FUSED) && (2 * ROWS * kUkfThreads * 8 + 2 * kUkfThreads * 4 <= 48 * 1024);
    __shared__ __align__(128) double s_in[PIPE ? 2 : 1][PIPE ? ROWS : 1][PIPE ? kUkfThreads : 1];
    __shared__ __align__(128) int s_st[PIPE ? 2 : 1][PIPE ? kUkfThreads : 1];
    __shared__ __align__(8) unsigned long long s_bar[2];
    const long long B = p.B;
    const int tid = threadIdx.x;
    const long long ntiles = (B + kUkfThreads - 1) / kUkfThreads;

    // stage the inputs of `tile` for this thread's filter into buffer `slot`
    auto prefetch = [&](long long tile, int slot) {
        const long long bb = (p.reverse ? ntiles - 1 - tile : tile) * kUkfThreads + tid;
        if (bb < B) {
#pragma unroll
            for (int r = 0; r < N; ++r) cp_async8(&s_in[slot][r][tid], p.x + (long long)r * B + bb);
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) cp_async8(&s_in[slot][N + r * (r + 1) / 2 + c][tid], p.P + (long long)(r * N + c) * B + bb);
#pragma unroll
            for (int c = 0; c < O; ++c) cp_async8(&s_in[slot][N + NTRI + c][tid], p.z + (long long)c * B + bb);
            cp_async4(&s_st[slot][tid], p.status + bb);
        }
    };
    // the same with TMA: thread 0 issues one bulk copy per SoA row of the tile (1 KB each, 512 B for the status row) that
    // complete on the slot's mbarrier — 18 instructions of one thread instead of 18 LDGSTS of every thread
    const bool tma = PIPE && p.use_tma != 0;
    auto prefetch_tma = [&](long long tile, int slot) {
        const long long b0 = (p.reverse ? ntiles - 1 - tile : tile) * kUkfThreads;
        const unsigned int cnt = (unsigned int)((B - b0 < kUkfThreads) ? (B - b0) : kUkfThreads);
        ukf_mbar_expect_tx(&s_bar[slot], cnt * 8u * ROWS + cnt * 4u);
#pragma unroll
        for (int r = 0; r < N; ++r) ukf_bulk_g2s(&s_in[slot][r][0], p.x + (long long)r * B + b0, cnt * 8u, &s_bar[slot]);
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c)
                ukf_bulk_g2s(&s_in[slot][N + r * (r + 1) / 2 + c][0], p.P + (long long)(r * N + c) * B + b0, cnt * 8u, &s_bar[slot]);
#pragma unroll
        for (int c = 0; c < O; ++c) ukf_bulk_g2s(&s_in[slot][N + NTRI + c][0], p.z + (long long)c * B + b0, cnt * 8u, &s_bar[slot]);
        ukf_bulk_g2s(&s_st[slot][0], p.status + b0, cnt * 4u, &s_bar[slot]);
    };
    if constexpr (PIPE) {
        if (tma) {
            if (tid == 0) {
                ukf_mbar_init(&s_bar[0], 1);
                ukf_mbar_init(&s_bar[1], 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
            if (tid == 0 && (long long)blockIdx.x < ntiles) prefetch_tma(blockIdx.x, 0);
        } else {
            if ((long long)blockIdx.x < ntiles) prefetch(blockIdx.x, 0);
            cp_async_commit();
        }
    }

  int it = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const long long b = (p.reverse ? ntiles - 1 - tile : tile) * kUkfThreads + tid;
    const bool live = b < B;
    double x[N], P[N][N], sig[N][M], z0[O];
    int st = MPCB_OK;
    if constexpr (PIPE) {
        const int slot = it & 1;
        if (tma) {
            ukf_mbar_wait(&s_bar[slot], (unsigned int)(it >> 1) & 1u);  // use number it/2 of this slot
        } else {
            if (tile + gridDim.x < ntiles) prefetch(tile + gridDim.x, slot ^ 1);
            cp_async_commit();   // one group per iteration (possibly empty): wait<1> always means "this tile has landed"
            cp_async_wait<1>();
        }
        if (live) {
            st = s_st[slot][tid];
#pragma unroll
            for (int r = 0; r < N; ++r) x[r] = s_in[slot][r][tid];
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c = 0; c < N; ++c) P[r][c] = (c <= r) ? s_in[slot][N + r * (r + 1) / 2 + c][tid] : 0.0;
#pragma unroll
            for (int c = 0; c < O; ++c) z0[c] = s_in[slot][N + NTRI + c][tid];
        }
        if (tma) {
            // every thread has taken its values out of `slot` (and, one iteration ago, out of the other buffer): only now may
            // thread 0 let the TMA overwrite the other buffer with the tile after this one
            __syncthreads();
            if (tid == 0 && tile + gridDim.x < ntiles) prefetch_tma(tile + gridDim.x, slot ^ 1);
        }
    } else if (live) {
        st = p.status[b];
#pragma unroll
        for (int r = 0; r < N; ++r) x[r] = p.x[(long long)r * B + b];
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) P[r][c] = (MODE == UKF_UPDATE || c <= r) ? p.P[(long long)(r * N + c) * B + b] : 0.0;
        if constexpr (MODE == UKF_UPDATE) {
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int i = 0; i < M; ++i) sig[r][i] = p.sigma_f[(long long)(r * M + i) * B + b];
        }
    }
    if (!live) continue;
    // x / P go back to memory only once they are fully defined in registers: a filter that enters with a sticky failure,
    // or whose first square root fails, leaves its state in memory as it was
    bool p_full = (MODE == UKF_UPDATE);

    const int steps = (MODE == UKF_FUSED) ? p.steps : 1;
    // the measurement of step s is fetched one step ahead (registers), so that the multi-step loop never waits on it
    double zc[O];
    if constexpr (MODE != UKF_PREDICT) {
#pragma unroll
        for (int c = 0; c < O; ++c) zc[c] = PIPE ? z0[c] : p.z[(long long)c * B + b];
    }
    for (int s = 0; s < steps && st == MPCB_OK; ++s) {
        double zn[O];
        if constexpr (MODE == UKF_FUSED) {
#pragma unroll
            for (int c = 0; c < O; ++c) zn[c] = (s + 1 < steps) ? p.z[((long long)(s + 1) * O + c) * B + b] : 0.0;
        }
        if constexpr (MODE != UKF_UPDATE) {
            // ---- predict (src/ukf.rs:44-52) ----
            const double u = p.has_u ? p.u[(long long)s * B + b] : p.u_scalar;
            double Lm[N][N];
            bool ok = true;
            if constexpr (SQRT == MPCB_SQRT_CHOLESKY) {
#pragma unroll
                for (int r = 0; r < N; ++r)
#pragma unroll
                    for (int c = 0; c < N; ++c) Lm[r][c] = p.cC * P[r][c];
                ok = chol_lower<N, FAST>(Lm);
            } else {
                double cp[N][N];
#pragma unroll
                for (int r = 0; r < N; ++r)
#pragma unroll
                    for (int c = 0; c < N; ++c) cp[r][c] = p.cC * P[r][c];
                sym_eig_sqrt<N, FAST>(cp, Lm);
            }
            if (!ok) { st = MPCB_CHOLESKY_FAIL; break; }
#pragma unroll
            for (int r = 0; r < N; ++r) {
                sig[r][0] = x[r];
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    sig[r][col_plus<N, ORDER>(i)] = x[r] + Lm[r][i];
                    sig[r][col_minus<N, ORDER>(i)] = x[r] - Lm[r][i];
                }
            }
#pragma unroll
            for (int i = 0; i < M; ++i) {
                double col[N];
#pragma unroll
                for (int r = 0; r < N; ++r) col[r] = sig[r][i];
                if constexpr (N >= 6) ukf_fx_call<MODEL, N, FAST>(p.mc, col, u, p.dt);
                else ukf_fx<MODEL, N, FAST>(p.mc, col, u, p.dt);
#pragma unroll
                for (int r = 0; r < N; ++r) sig[r][i] = col[r];
            }
            unscented_transform<N, M, FAST>(sig, p.wm0, p.wc0, p.wi, p.Q, x, P);
            p_full = true;
        }
        if constexpr (MODE != UKF_PREDICT) {
            // ---- update (src/ukf.rs:54-74) ----
            double zs[O][M];
#pragma unroll
            for (int i = 0; i < M; ++i) {
                double col[N], zz[O];
#pragma unroll
                for (int r = 0; r < N; ++r) col[r] = sig[r][i];
                if constexpr (N >= 6) ukf_hx_call<MODEL, N, O, FAST>(p.mc, col, zz);
                else ukf_hx<MODEL, N, O, FAST>(p.mc, col, zz);
#pragma unroll
                for (int r = 0; r < O; ++r) zs[r][i] = ((p.enable >> r) & 1u) ? zz[r] : 0.0;
            }
            double zp[O], pz[O][O];
            unscented_transform<O, M, FAST>(zs, p.wm0, p.wc0, p.wi, p.R, zp, pz);
            double pxz[N][O];
            if constexpr (!FAST) {
#pragma unroll
                for (int r = 0; r < N; ++r)
#pragma unroll
                    for (int c = 0; c < O; ++c) pxz[r][c] = 0.0;
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    const double w = (i == 0) ? p.wc0 : p.wi;
#pragma unroll
                    for (int r = 0; r < N; ++r) {
                        const double wdx = w * (sig[r][i] - x[r]);
#pragma unroll
                        for (int c = 0; c < O; ++c) pxz[r][c] += wdx * (zs[c][i] - zp[c]);
                    }
                }
            } else {
                // sum_{i>=1} dx_i dz_i^T first (shared weight factored out), then the i = 0 term
#pragma unroll
                for (int r = 0; r < N; ++r)
#pragma unroll
                    for (int c = 0; c < O; ++c) pxz[r][c] = 0.0;
#pragma unroll
                for (int i = 1; i < M; ++i) {
                    double dz[O];
#pragma unroll
                    for (int c = 0; c < O; ++c) dz[c] = zs[c][i] - zp[c];
#pragma unroll
                    for (int r = 0; r < N; ++r) {
                        const double dx = sig[r][i] - x[r];
#pragma unroll
                        for (int c = 0; c < O; ++c) pxz[r][c] = fma(dx, dz[c], pxz[r][c]);
                    }
                }
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    const double wdx = p.wc0 * (sig[r][0] - x[r]);
#pragma unroll
                    for (int c = 0; c < O; ++c) pxz[r][c] = fma(p.wi, pxz[r][c], wdx * (zs[c][0] - zp[c]));
                }
            }
            double pzi[O][O];
            if (!inverse_small<O, FAST>(pz, pzi)) { st = MPCB_INVERSE_FAIL; break; }
            double k[N][O];
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c = 0; c < O; ++c) {
                    double acc = pxz[r][0] * pzi[0][c];
#pragma unroll
                    for (int j = 1; j < O; ++j) acc += pxz[r][j] * pzi[j][c];
                    k[r][c] = acc;
                }
            double innov[O];
#pragma unroll
            for (int c = 0; c < O; ++c) innov[c] = zc[c] - zp[c];
#pragma unroll
            for (int r = 0; r < N; ++r) {
                double acc = k[r][0] * innov[0];
#pragma unroll
                for (int j = 1; j < O; ++j) acc += k[r][j] * innov[j];
                x[r] += acc;
            }
            double kp[N][O];
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c = 0; c < O; ++c) {
                    double acc = k[r][0] * pz[0][c];
#pragma unroll
                    for (int j = 1; j < O; ++j) acc += k[r][j] * pz[j][c];
                    kp[r][c] = acc;
                }
            if constexpr (!FAST) {
#pragma unroll
                for (int r = 0; r < N; ++r)
#pragma unroll
                    for (int c = 0; c < N; ++c) {
                        double acc = kp[r][0] * k[c][0];
#pragma unroll
                        for (int j = 1; j < O; ++j) acc += kp[r][j] * k[c][j];
                        P[r][c] -= acc;
                    }
#pragma unroll
                for (int r = 0; r < N; ++r)
#pragma unroll
                    for (int c = r; c < N; ++c) {
                        const double a = (P[r][c] + P[c][r]) / 2.0;
                        P[r][c] = a;
                        P[c][r] = a;
                    }
            } else {
                // K Pz K^T is symmetric and P is kept exactly symmetric by the fast transform: upper triangle only
#pragma unroll
                for (int r = 0; r < N; ++r)
#pragma unroll
                    for (int c = r; c < N; ++c) {
                        double acc = kp[r][0] * k[c][0];
#pragma unroll
                        for (int j = 1; j < O; ++j) acc = fma(kp[r][j], k[c][j], acc);
                        const double a = P[r][c] - acc;
                        P[r][c] = a;
                        P[c][r] = a;
                    }
            }
        }
        if constexpr (MODE == UKF_FUSED) {
#pragma unroll
            for (int c = 0; c < O; ++c) zc[c] = zn[c];
        }
    }

    if (p_full) {
#pragma unroll
        for (int r = 0; r < N; ++r) p.x[(long long)r * B + b] = x[r];
        // after an update P is exactly symmetric (src/ukf.rs:73 symmetrises it): with lower_only the strictly-upper
        // triangle stays in memory as it was — predict never reads it, and the read-out calls mirror the lower one
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c)
                if (c <= r || !(MODE == UKF_FUSED && p.lower_only)) p.P[(long long)(r * N + c) * B + b] = P[r][c];
    }
    if constexpr (MODE == UKF_PREDICT) {
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int i = 0; i < M; ++i) p.sigma_f[(long long)(r * M + i) * B + b] = sig[r][i];
    }
    p.status[b] = st;
  }
  if constexpr (PIPE) cp_async_wait<0>();
}

// AoS <-> SoA transposes for the host-facing calls: in[B][W] <-> out[W][B]
__global__ void ukf_aos_to_soa(const double* __restrict__ in, double* __restrict__ out, long long B, int W);
__global__ void ukf_soa_to_aos(const double* __restrict__ in, double* __restrict__ out, long long B, int W,
                               long long first, long long count);
__global__ void ukf_broadcast(const double* __restrict__ row, double* __restrict__ out, long long B, int W);

using UkfKernelFn = void (*)(const UkfParams);
// exact: reference operation order, no FMA (ukf_n4.cu / ukf_n6.cu, -fmad=false); fast: see unscented_transform
UkfKernelFn ukf_kernel_n4(int model_id, int sqrt_mode, int order, int mode);
UkfKernelFn ukf_kernel_n6(int model_id, int sqrt_mode, int order, int mode);
UkfKernelFn ukf_kernel_n4_fast(int model_id, int sqrt_mode, int order, int mode);
UkfKernelFn ukf_kernel_n6_fast(int model_id, int sqrt_mode, int order, int mode);
// fused fast kernels of the six-state models in the streaming form (ukf_stream_kernel.cuh, ukf_n6_stream.cu)
UkfKernelFn ukf_stream_kernel_n6(int model_id, int sqrt_mode, int order, size_t* smem_bytes);

}  // namespace mpcb
