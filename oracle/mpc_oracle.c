/*
 * mpc_oracle.c — CPU restatement of the mpc-rs MPPI / UKF hot path.  TEST INFRASTRUCTURE ONLY;
 * PARITY UNPINNED (the reference ships no tests or golden vectors) — see mpc_oracle.h.
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp; no -ffast-math).
 */
#include "mpc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI 3.14159265358979323846264338327950288 /* std::f64::consts::PI */
#define ORC_MAX_N 6
#define ORC_MAX_M (2 * ORC_MAX_N + 1)

/* ------------------------------------------------------------------ model constants ---- */

/* Constants of each example, evaluated in the example's own order. */
int orc_model_defaults(int model_id, mpcb_model_params* p) {
    memset(p, 0, sizeof(*p));
    switch (model_id) {
        case MPCB_MODEL_L:       /* examples/mppi4.rs:8-18,73-81 */
        case MPCB_MODEL_NL:      /* examples/mppi4-non-liner.rs:8-18,73-80 */
        case MPCB_MODEL_PEN_LIN: /* examples/ukf-pen.rs:6-16 */
        case MPCB_MODEL_PEN_NL:  /* examples/ukf-pen2.rs:8-17 */
        case MPCB_MODEL_PEN6:    /* examples/ukf-pen3.rs:8-17 */
            p->m1 = 150e-3;
            p->r_w = 50e-3;
            p->m2 = 2.3 - 2.0 * p->m1 + 2.0;
            p->l = 0.2474;
            p->j1 = p->m1 * p->r_w * p->r_w;
            p->j2 = (model_id == MPCB_MODEL_PEN_LIN) ? 0.1 : 0.2;
            p->g = 9.81;
            p->kt = 0.15;
            if (model_id == MPCB_MODEL_L || model_id == MPCB_MODEL_NL) {
                p->dt = 0.8 / (double)8; /* T / N */
                p->cost[0] = 2.0; p->cost[1] = 3.0; p->cost[2] = 5.0; p->cost[3] = 1.2;
                p->cost[4] = 2.0; p->cost[5] = 5.0; p->cost[6] = 2.0; p->cost[7] = 0.35; p->cost[8] = 0.75;
            } else {
                p->dt = 0.01;
            }
            return MPCB_OK;
        case MPCB_MODEL_NL6:     /* examples/mppi4-non-liner-ukf.rs:13-24,108-124 */
        case MPCB_MODEL_NL6_UKF:
            p->m1 = 160e-3;
            p->r_w = 50e-3;
            p->m2 = 2.4;
            p->l = 0.4;
            p->j1 = 2.23e5 * 1e-9;
            p->j2 = 1.168e8 * 1e-9;
            p->g = 9.81;
            p->kt = 0.15;
            p->dt = 1.2 / (double)8;
            p->cost[0] = 0.1; p->cost[1] = 0.1; p->cost[2] = 1.0; p->cost[3] = 0.5;
            return MPCB_OK;
        default: return MPCB_BAD_ARG;
    }
}

int orc_model_dims(int model_id, int* n, int* o) {
    switch (model_id) {
        case MPCB_MODEL_PEN_LIN: *n = 4; *o = 2; return MPCB_OK;
        case MPCB_MODEL_PEN_NL: *n = 4; *o = 3; return MPCB_OK;
        case MPCB_MODEL_PEN6:
        case MPCB_MODEL_NL6_UKF: *n = 6; *o = 5; return MPCB_OK;
        default: return MPCB_BAD_ARG;
    }
}

/* ------------------------------------------------------------ MPPI models, f64 + f32 ---- */
#define REAL double
#define FN(name) name##_d
#define R_SIN sin
#define R_COS cos
#include "models_impl.inc"
#undef REAL
#undef FN
#undef R_SIN
#undef R_COS

#define REAL float
#define FN(name) name##_f
#define R_SIN sinf
#define R_COS cosf
#include "models_impl.inc"
#undef REAL
#undef FN
#undef R_SIN
#undef R_COS

void orc_dynamics(int model_id, const mpcb_model_params* p, const double* x, double u, double* out) {
    double tmp[4];
    dynamics_d(model_id, p, x, u, tmp);
    memcpy(out, tmp, sizeof(tmp));
}

double orc_cost(int model_id, const mpcb_model_params* p, const double* x) { return cost_d(model_id, p, x); }

void orc_ddot(const mpcb_model_params* p, const double* x4, double u, double f, double* ddx, double* ddth) {
    ddot_d(p, x4, u, f, ddx, ddth);
}

/* examples/mppi4-non-liner-ukf.rs:149-159 */
void orc_dynamics_short(const mpcb_model_params* p, const double* x, double u, double dt, double f, double* out) {
    double x4[4] = {x[0], x[1], x[3], x[4]};
    double ddx, ddth;
    ddot_d(p, x4, u, f, &ddx, &ddth);
    double r[6];
    memcpy(r, x, sizeof(r));
    r[5] = ddth;
    r[4] += r[5] * dt;
    r[3] += r[4] * dt;
    r[2] = ddx;
    r[1] += r[2] * dt;
    r[0] += r[1] * dt;
    memcpy(out, r, sizeof(r));
}

/* ------------------------------------------------------------------ UKF models ---- */

/* examples/ukf-pen3.rs:35-50 — d uses x[2].cos() although theta is x[3] in this layout (as written) */
static void fx_pen6(const mpcb_model_params* p, const double* x, double u, double dtv, double* r) {
    const double M1 = p->m1, R_W = p->r_w, M2 = p->m2, L = p->l, J1 = p->j1, J2 = p->j2, G = p->g, KT = p->kt;
    const double D = (M1 + M2 + J1 / (R_W * R_W)) * (M2 * L * L + J2);
    const double mlc = M2 * L * cos(x[2]);
    double d = D - mlc * mlc;
    double o[6];
    memcpy(o, x, sizeof(o));
    o[0] += x[1] * dtv;
    o[1] += x[2] * dtv;
    double term3 = (J2 + M2 * L * L) * (KT * u / R_W + M2 * L * (x[4] * x[4]) * sin(x[3]));
    double term4 = M2 * G * L * L * sin(x[3]) * cos(x[3]);
    o[2] = (term3 + term4) / d;
    o[3] += x[4] * dtv;
    o[4] += x[5] * dtv;
    double term1 = (M1 + M2 + J1 / (R_W * R_W)) * M2 * G * L * sin(x[3]);
    double term2 = (KT * u / R_W + M2 * L * (x[4] * x[4]) * sin(x[3])) * M2 * L * cos(x[3]);
    o[5] = (term1 - term2) / d;
    memcpy(r, o, sizeof(o));
}

static double to_degrees(double x) { return x * (180.0 / ORC_PI); } /* f64::to_degrees */

void orc_fx(int model_id, const mpcb_model_params* p, const double* x, double u, double dt, double* out) {
    const double dtv = (dt > 0.0) ? dt : p->dt;
    double tmp[6];
    switch (model_id) {
        case MPCB_MODEL_PEN_LIN: dyn_linear_d(p, x, u, dtv, tmp); memcpy(out, tmp, 4 * sizeof(double)); break;
        case MPCB_MODEL_PEN_NL: dyn_nonlinear_d(p, x, u, dtv, tmp); memcpy(out, tmp, 4 * sizeof(double)); break;
        case MPCB_MODEL_PEN6: fx_pen6(p, x, u, dtv, out); break;
        case MPCB_MODEL_NL6_UKF: orc_dynamics_short(p, x, u, dtv, 0.0, out); break;
        default: break;
    }
}

void orc_hx(int model_id, const mpcb_model_params* p, const double* x, double* z) {
    const double R_W = p->r_w, M2 = p->m2, L = p->l, G = p->g;
    switch (model_id) {
        case MPCB_MODEL_PEN_LIN: /* examples/ukf-pen.rs:86-91 */
            z[0] = x[1];
            z[1] = x[3];
            break;
        case MPCB_MODEL_PEN_NL: /* examples/ukf-pen2.rs:47-53 */
            z[0] = 60.0 / (2.0 * ORC_PI * R_W) * x[1];
            z[1] = 60.0 / (2.0 * ORC_PI * R_W) * x[1];
            z[2] = to_degrees(x[3]);
            break;
        case MPCB_MODEL_PEN6: { /* examples/ukf-pen3.rs:53-63 */
            double v = M2 * G * cos(x[3]) + M2 * x[2] * sin(x[3]) - M2 * L * (x[4] * x[4]);
            double h = -M2 * G * sin(x[3]) + M2 * x[2] * cos(x[3]) + M2 * L * x[5];
            z[0] = 60.0 / (2.0 * ORC_PI * R_W) * x[1];
            z[1] = 60.0 / (2.0 * ORC_PI * R_W) * x[1];
            z[2] = to_degrees(x[3]);
            z[3] = v / G;
            z[4] = h / G;
            break;
        }
        case MPCB_MODEL_NL6_UKF: { /* examples/mppi4-non-liner-ukf.rs:169-179 */
            double ax = G * sin(x[3]) + x[2] * cos(x[3]) + L * x[5];
            double az = G * cos(x[3]) - x[2] * sin(x[3]) + L * (x[4] * x[4]);
            z[0] = 36.0 * 60.0 / (2.0 * ORC_PI * R_W) * x[1];
            z[1] = 36.0 * -60.0 / (2.0 * ORC_PI * R_W) * x[1];
            z[2] = to_degrees(x[4]);
            z[3] = az / G;
            z[4] = ax / G;
            break;
        }
        default: break;
    }
}

/* examples/mppi4-non-liner-ukf.rs:192-221 with PHY = (100, 70, 20) (:27) */
void orc_gen_q(double dt, double* q) {
    const double phy0 = 100.0, phy1 = 70.0, phy2 = 20.0;
    double dt_2 = dt * dt;
    double dt_3 = dt_2 * dt;
    double dt_4 = dt_2 * dt_2;
    double q1[36] = {0}, q2[36] = {0}, q3[36] = {0};
#define AT(m, r, c) m[(r) * 6 + (c)]
    AT(q1, 3, 4) = dt_4 / 8.0; AT(q1, 3, 5) = dt_3 / 6.0;
    AT(q1, 4, 3) = dt_4 / 8.0; AT(q1, 4, 4) = dt_3 / 3.0; AT(q1, 4, 5) = dt_2 / 2.0;
    AT(q1, 5, 3) = dt_3 / 6.0; AT(q1, 5, 4) = dt_2 / 2.0; AT(q1, 5, 5) = dt;
    AT(q2, 1, 3) = dt_4 / 8.0; AT(q2, 1, 4) = dt_3 / 6.0;
    AT(q2, 3, 1) = dt_4 / 8.0; AT(q2, 3, 3) = dt_3 / 3.0; AT(q2, 3, 4) = dt_2 / 2.0;
    AT(q2, 4, 1) = dt_3 / 6.0; AT(q2, 4, 3) = dt_2 / 2.0; AT(q2, 4, 4) = dt;
    AT(q3, 0, 1) = dt_4 / 8.0; AT(q3, 0, 2) = dt_3 / 6.0;
    AT(q3, 1, 0) = dt_4 / 8.0; AT(q3, 1, 1) = dt_3 / 3.0; AT(q3, 1, 2) = dt_2 / 2.0;
    AT(q3, 2, 0) = dt_3 / 6.0; AT(q3, 2, 1) = dt_2 / 2.0; AT(q3, 2, 2) = dt;
#undef AT
    for (int i = 0; i < 36; ++i) q[i] = phy0 * q1[i] + phy1 * q2[i] + phy2 * q3[i];
}

int orc_ukf_default_noise(int model_id, double dt, double* Q, double* R, double* P0) {
    int n, o;
    if (orc_model_dims(model_id, &n, &o) != MPCB_OK) return MPCB_BAD_ARG;
    memset(Q, 0, sizeof(double) * n * n);
    memset(R, 0, sizeof(double) * o * o);
    memset(P0, 0, sizeof(double) * n * n);
    for (int i = 0; i < n; ++i) P0[i * n + i] = 10.0;
    switch (model_id) {
        case MPCB_MODEL_PEN_LIN: /* examples/ukf-pen.rs:17-26,148-153 */
            Q[1 * 4 + 1] = 1.0; Q[2 * 4 + 2] = 0.25; Q[2 * 4 + 3] = 0.5; Q[3 * 4 + 2] = 0.5; Q[3 * 4 + 3] = 1.0;
            R[0] = 0.5; R[3] = 0.5;
            break;
        case MPCB_MODEL_PEN_NL: /* examples/ukf-pen2.rs:18-28,71-76 */
            Q[3 * 4 + 3] = 0.25;
            R[0] = 100.0; R[4] = 100.0; R[8] = 0.5;
            break;
        case MPCB_MODEL_PEN6: /* examples/ukf-pen3.rs:18-32,83-90 */
            Q[5 * 6 + 5] = 10.0;
            R[0] = 100.0; R[6] = 100.0; R[12] = 0.5; R[18] = 100.0; R[24] = 100.0;
            break;
        case MPCB_MODEL_NL6_UKF: { /* examples/mppi4-non-liner-ukf.rs:28,161-167 */
            const double rdiag[5] = {200.0, 200.0, 10.0, 0.05, 0.05};
            orc_gen_q(dt > 0.0 ? dt : 1.2 / 8.0, Q);
            for (int i = 0; i < 5; ++i) R[i * 5 + i] = rdiag[i];
            break;
        }
        default: return MPCB_BAD_ARG;
    }
    return MPCB_OK;
}

/* ------------------------------------------------------------------ MPPI ---- */

static int is_finite_d(double v) { return isfinite(v); }

/* PASS 3-6 of src/mppi.rs:65-91 on c[K] and v[K][H] (v element type given by vget). */
static int mppi_finish(int64_t K, int H, double lambda, double* w /* in: c_k, out: w_k */, const void* v,
                       int v_is_float, double* u_out, orc_mppi_out* info) {
    /* :65-69 finite-only max */
    double mx = 0.0;
    int64_t arg = -1, nfin = 0;
    for (int64_t k = 0; k < K; ++k) {
        if (is_finite_d(w[k])) {
            ++nfin;
            if (arg < 0 || w[k] > mx) { mx = w[k]; arg = k; }
        }
    }
    if (info) { info->argmax = arg; info->n_finite = nfin; info->max = mx; info->sum = 0.0; }
    if (arg < 0) {
        if (info) info->status = MPCB_NO_FINITE_COST;
        return MPCB_NO_FINITE_COST;
    }
    /* :71-74 */
    for (int64_t k = 0; k < K; ++k) w[k] = exp((w[k] - mx) / lambda);
    double sum = 0.0;
    for (int64_t k = 0; k < K; ++k) sum += w[k];
    if (info) info->sum = sum;
    if (sum == 0.0) {
        if (info) info->status = MPCB_SUM_ZERO;
        return MPCB_SUM_ZERO;
    }
    /* :80-84 — sequential order (the reference's rayon tree order is unspecified) */
    for (int t = 0; t < H; ++t) u_out[t] = 0.0;
    for (int64_t k = 0; k < K; ++k) {
        double wn = w[k] / sum;
        if (v_is_float) {
            const float* vk = (const float*)v + k * (int64_t)H;
            for (int t = 0; t < H; ++t) u_out[t] += wn * (double)vk[t];
        } else {
            const double* vk = (const double*)v + k * (int64_t)H;
            for (int t = 0; t < H; ++t) u_out[t] += wn * vk[t];
        }
    }
    /* :87-89 only element 0 is checked */
    if (isnan(u_out[0]) || isinf(u_out[0])) {
        if (info) info->status = MPCB_U_INVALID;
        return MPCB_U_INVALID;
    }
    if (info) info->status = MPCB_OK;
    return MPCB_OK;
}

int orc_mppi_compute(int model_id, const mpcb_model_params* p, int64_t K, int H, double lambda, double std_dev,
                     double lo, double hi, const double* x, const double* u_n, const double* eps,
                     double* u_out, double* c_out, orc_mppi_out* info) {
    if (K <= 0 || H <= 0) return MPCB_BAD_ARG;
    double* v = (double*)malloc(sizeof(double) * (size_t)K * (size_t)H);
    double* c = (double*)malloc(sizeof(double) * (size_t)K);
    if (!v || !c) { free(v); free(c); return MPCB_BAD_ARG; }
    /* :39-45 v = clamp(u_n + eps) */
    for (int64_t k = 0; k < K; ++k)
        for (int t = 0; t < H; ++t) v[k * H + t] = clampr_d(u_n[t] + eps[k * H + t], lo, hi);
    /* :48 */
    const double inv = 1.0 / (std_dev * std_dev); /* powi(-2) */
    /* :49-63 */
    for (int64_t k = 0; k < K; ++k) {
        const double* vk = v + k * H;
        double xc[4] = {x[0], x[1], x[2], x[3]}, xn[4];
        double cost = 0.0;
        for (int t = 0; t < H; ++t) {
            dynamics_d(model_id, p, xc, vk[t], xn);
            cost = cost + cost_d(model_id, p, xn);
            memcpy(xc, xn, sizeof(xc));
        }
        double control_term = 0.0;
        for (int t = 0; t < H; ++t) control_term += u_n[t] * inv * vk[t];
        c[k] = -cost - control_term;
    }
    if (c_out) memcpy(c_out, c, sizeof(double) * (size_t)K);
    int st = mppi_finish(K, H, lambda, c, v, 0, u_out, info);
    free(v);
    free(c);
    return st;
}

int orc_mppi_compute_f32(int model_id, const mpcb_model_params* p, int64_t K, int H, double lambda, double std_dev,
                         double lo, double hi, const double* x, const double* u_n, const float* eps,
                         double* u_out, double* c_out, orc_mppi_out* info) {
    if (K <= 0 || H <= 0) return MPCB_BAD_ARG;
    float* v = (float*)malloc(sizeof(float) * (size_t)K * (size_t)H);
    double* c = (double*)malloc(sizeof(double) * (size_t)K);
    if (!v || !c) { free(v); free(c); return MPCB_BAD_ARG; }
    const float flo = (float)lo, fhi = (float)hi;
    for (int64_t k = 0; k < K; ++k)
        for (int t = 0; t < H; ++t) v[k * H + t] = clampr_f((float)u_n[t] + eps[k * H + t], flo, fhi);
    const double inv = 1.0 / (std_dev * std_dev);
    for (int64_t k = 0; k < K; ++k) {
        const float* vk = v + k * H;
        float xc[4] = {(float)x[0], (float)x[1], (float)x[2], (float)x[3]}, xn[4];
        double cost = 0.0, control_term = 0.0;
        for (int t = 0; t < H; ++t) {
            dynamics_f(model_id, p, xc, vk[t], xn);
            cost = cost + (double)cost_f(model_id, p, xn);
            memcpy(xc, xn, sizeof(xc));
            control_term += (u_n[t] * inv) * (double)vk[t];
        }
        c[k] = -cost - control_term;
    }
    if (c_out) memcpy(c_out, c, sizeof(double) * (size_t)K);
    int st = mppi_finish(K, H, lambda, c, v, 1, u_out, info);
    free(v);
    free(c);
    return st;
}

/* ------------------------------------------- CPU baseline: RNG of the reference ---- */
/* rand_xoshiro 0.6.0 Xoshiro256Plus (xoshiro256+ 1.0, Blackman & Vigna), SplitMix64 seeding like
 * SeedableRng::seed_from_u64; the reference seeds from OS entropy (src/mppi.rs:41). */
typedef struct { uint64_t s[4]; } xoshiro256p;

static uint64_t splitmix64(uint64_t* st) {
    uint64_t z = (*st += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static void xo_seed(xoshiro256p* g, uint64_t seed) {
    for (int i = 0; i < 4; ++i) g->s[i] = splitmix64(&seed);
}
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t xo_next(xoshiro256p* g) {
    const uint64_t result = g->s[0] + g->s[3];
    const uint64_t t = g->s[1] << 17;
    g->s[2] ^= g->s[0];
    g->s[3] ^= g->s[1];
    g->s[1] ^= g->s[2];
    g->s[0] ^= g->s[3];
    g->s[2] ^= t;
    g->s[3] = rotl64(g->s[3], 45);
    return result;
}

/* 256-layer ziggurat for N(0,1) (Marsaglia & Tsang 2000 as refined by Doornik 2005; rand_distr 0.4.3's
 * StandardNormal is a 256-layer ziggurat with R = 3.654152885361009).  Tables are computed here at
 * first use because the crate's tables are not in this image. */
#define ZIG_N 256
#define ZIG_R 3.654152885361009
#define ZIG_V 0.00492867323399 /* area of each layer for N=256 */
static double zig_x[ZIG_N + 1], zig_f[ZIG_N + 1];
static int zig_ready = 0;
static void zig_init(void) {
    if (zig_ready) return;
    double f = exp(-0.5 * ZIG_R * ZIG_R);
    zig_x[0] = ZIG_V / f; /* base strip pseudo-width */
    zig_x[1] = ZIG_R;
    zig_x[ZIG_N] = 0.0;
    for (int i = 2; i < ZIG_N; ++i) {
        zig_x[i] = sqrt(-2.0 * log(ZIG_V / zig_x[i - 1] + f));
        f = exp(-0.5 * zig_x[i] * zig_x[i]);
    }
    for (int i = 0; i <= ZIG_N; ++i) zig_f[i] = exp(-0.5 * zig_x[i] * zig_x[i]);
    zig_ready = 1;
}
static inline double u01_open(xoshiro256p* g) { /* (0,1) */
    return ((double)(xo_next(g) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
static double zig_normal(xoshiro256p* g) {
    for (;;) {
        uint64_t bits = xo_next(g);
        int i = (int)(bits & 0xff);
        /* symmetric uniform in (-1,1) from the high 53 bits */
        double u = 2.0 * (((double)(bits >> 11) + 0.5) * (1.0 / 9007199254740992.0)) - 1.0;
        double xv = u * zig_x[i];
        if (fabs(xv) < zig_x[i + 1]) return xv; /* inside the rectangle core */
        if (i == 0) { /* tail */
            double a, b;
            do {
                a = log(u01_open(g)) / ZIG_R;
                b = log(u01_open(g));
            } while (-2.0 * b < a * a);
            return (u < 0.0) ? a - ZIG_R : ZIG_R - a;
        }
        /* wedge */
        if (zig_f[i + 1] + (zig_f[i] - zig_f[i + 1]) * u01_open(g) < exp(-0.5 * xv * xv)) return xv;
    }
}

void orc_normal_fill(uint64_t seed, double* out, int64_t n) {
    zig_init();
    xoshiro256p g;
    xo_seed(&g, seed);
    for (int64_t i = 0; i < n; ++i) out[i] = zig_normal(&g);
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* The reference's structure (src/mppi.rs:38-84): six passes over materialised v[K][H], each a
 * parallel loop over samples; one RNG per worker (map_init, :41). */
int orc_mppi_compute_cpu(int model_id, const mpcb_model_params* p, int64_t K, int H, double lambda, double std_dev,
                         double lo, double hi, const double* x, const double* u_n, uint64_t seed, int threads,
                         double* u_out, orc_mppi_out* info) {
    if (K <= 0 || H <= 0 || H > 4096) return MPCB_BAD_ARG;
    zig_init();
    if (threads <= 0) threads = orc_max_threads();
    double* v = (double*)malloc(sizeof(double) * (size_t)K * (size_t)H);
    double* c = (double*)malloc(sizeof(double) * (size_t)K);
    if (!v || !c) { free(v); free(c); return MPCB_BAD_ARG; }
    /* PASS 1 :39-45 */
#pragma omp parallel num_threads(threads)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        xoshiro256p g;
        xo_seed(&g, seed * 0x9E3779B97F4A7C15ull + (uint64_t)tid * 0xD1B54A32D192ED03ull + 1);
#pragma omp for schedule(static)
        for (int64_t k = 0; k < K; ++k)
            for (int t = 0; t < H; ++t) v[k * H + t] = clampr_d(u_n[t] + std_dev * zig_normal(&g), lo, hi);
    }
    /* PASS 2 :48-63 */
    const double inv = 1.0 / (std_dev * std_dev);
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int64_t k = 0; k < K; ++k) {
        const double* vk = v + k * H;
        double xc[4] = {x[0], x[1], x[2], x[3]}, xn[4];
        double cost = 0.0;
        for (int t = 0; t < H; ++t) {
            dynamics_d(model_id, p, xc, vk[t], xn);
            cost = cost + cost_d(model_id, p, xn);
            memcpy(xc, xn, sizeof(xc));
        }
        double control_term = 0.0;
        for (int t = 0; t < H; ++t) control_term += u_n[t] * inv * vk[t];
        c[k] = -cost - control_term;
    }
    /* PASS 3 :65-69 */
    double mx = -INFINITY;
    int64_t nfin = 0;
#pragma omp parallel for reduction(max : mx) reduction(+ : nfin) schedule(static) num_threads(threads)
    for (int64_t k = 0; k < K; ++k)
        if (isfinite(c[k])) { ++nfin; if (c[k] > mx) mx = c[k]; }
    int64_t arg = -1;
    if (nfin > 0)
        for (int64_t k = 0; k < K; ++k)
            if (c[k] == mx) { arg = k; break; }
    if (info) { info->argmax = arg; info->n_finite = nfin; info->max = mx; info->sum = 0.0; info->status = MPCB_OK; }
    int st = MPCB_OK;
    if (nfin == 0) {
        st = MPCB_NO_FINITE_COST;
    } else {
        /* PASS 4 :71-73 */
#pragma omp parallel for schedule(static) num_threads(threads)
        for (int64_t k = 0; k < K; ++k) c[k] = exp((c[k] - mx) / lambda);
        /* PASS 5 :74 */
        double sum = 0.0;
#pragma omp parallel for reduction(+ : sum) schedule(static) num_threads(threads)
        for (int64_t k = 0; k < K; ++k) sum += c[k];
        if (info) info->sum = sum;
        if (sum == 0.0) {
            st = MPCB_SUM_ZERO;
        } else {
            /* PASS 6 :80-84 */
            for (int t = 0; t < H; ++t) u_out[t] = 0.0;
#pragma omp parallel num_threads(threads)
            {
                double* acc = (double*)calloc((size_t)H, sizeof(double));
#pragma omp for schedule(static) nowait
                for (int64_t k = 0; k < K; ++k) {
                    const double wn = c[k] / sum;
                    const double* vk = v + k * H;
                    for (int t = 0; t < H; ++t) acc[t] += wn * vk[t];
                }
#pragma omp critical
                for (int t = 0; t < H; ++t) u_out[t] += acc[t];
                free(acc);
            }
            if (isnan(u_out[0]) || isinf(u_out[0])) st = MPCB_U_INVALID;
        }
    }
    if (info) info->status = st;
    free(v);
    free(c);
    return st;
}

/* ------------------------------------------------------------------ UKF ---- */

/* src/ukf.rs:23-28,112-118 (same constants in ukf2.rs and examples/ukf-pen.rs:28-42) */
void orc_ukf_weights(int n, double* wm, double* wc) {
    const double N = (double)n, ALPHA = 1e-3, BETA = 2.0;
    const double KAPPA = 3.0 - N;
    const double C = ALPHA * ALPHA * (N + KAPPA);
    const double LAMBDA = C - N;
    const int M = 2 * n + 1;
    for (int i = 0; i < M; ++i) { wm[i] = 1.0 / (2.0 * C); wc[i] = 1.0 / (2.0 * C); }
    wm[0] = LAMBDA / C;
    wc[0] = LAMBDA / C + 1.0 - ALPHA * ALPHA + BETA;
}

static double ukf_c(int n) {
    const double N = (double)n, ALPHA = 1e-3;
    const double KAPPA = 3.0 - N;
    return ALPHA * ALPHA * (N + KAPPA);
}

/* nalgebra 0.33.0 Cholesky (lower; reads the lower triangle only): column j is reduced by the previous
 * columns in order, the diagonal must be > 0, the sub-column is divided by its root. */
int orc_cholesky_lower(int n, const double* A, double* L) {
    double m[ORC_MAX_N * ORC_MAX_N];
    memcpy(m, A, sizeof(double) * n * n);
    for (int j = 0; j < n; ++j) {
        for (int k = 0; k < j; ++k) {
            const double ljk = m[j * n + k];
            for (int i = j; i < n; ++i) m[i * n + j] -= m[i * n + k] * ljk;
        }
        const double diag = m[j * n + j];
        if (!(diag > 0.0)) return MPCB_CHOLESKY_FAIL;
        const double denom = sqrt(diag);
        m[j * n + j] = denom;
        for (int i = j + 1; i < n; ++i) m[i * n + j] /= denom;
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) L[i * n + j] = (j <= i) ? m[i * n + j] : 0.0;
    return MPCB_OK;
}

/* U*sqrt(S) of svd_unordered(C*P) (src/ukf.rs:121-124).  For a symmetric PSD matrix the SVD is the
 * eigendecomposition; nalgebra's SVD routine is not available here, so this is the published cyclic
 * Jacobi method (Golub & Van Loan, Alg. 8.5.x: symmetric Schur rotations, row-cyclic sweeps, at most
 * ORC_JACOBI_SWEEPS sweeps) on the lower triangle.  Column sign/order of U are immaterial (±L_i carry equal weights). */
#define ORC_JACOBI_SWEEPS 10
void orc_sym_eig_sqrt(int n, const double* Ain, double* Lout) {
    double A[ORC_MAX_N][ORC_MAX_N], V[ORC_MAX_N][ORC_MAX_N];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            A[i][j] = (j <= i) ? Ain[i * n + j] : Ain[j * n + i];
            V[i][j] = (i == j) ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < ORC_JACOBI_SWEEPS; ++sweep) {
        double off = 0.0; /* stop once every off-diagonal element is exactly zero */
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) off += A[p][q] * A[p][q];
        if (off == 0.0) break;
        for (int p = 0; p < n - 1; ++p) {
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[p][q];
                if (apq == 0.0) continue;
                const double tau = (A[q][q] - A[p][p]) / (2.0 * apq);
                const double rt = sqrt(1.0 + tau * tau);
                const double t = (tau >= 0.0) ? 1.0 / (tau + rt) : -1.0 / (-tau + rt);
                const double c = 1.0 / sqrt(1.0 + t * t);
                const double s = t * c;
                for (int k = 0; k < n; ++k) { /* A <- A*J */
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) { /* A <- J^T*A */
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                A[p][q] = 0.0;
                A[q][p] = 0.0;
                for (int k = 0; k < n; ++k) { /* V <- V*J */
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
        }
    }
    for (int j = 0; j < n; ++j) {
        const double ev = A[j][j];
        const double sq = sqrt(fabs(ev)); /* singular value = |eigenvalue| */
        for (int i = 0; i < n; ++i) Lout[i * n + j] = V[i][j] * sq;
    }
}

/* nalgebra try_inverse: closed forms for 1x1..3x3 (determinant == 0 -> None); LU with partial
 * pivoting for larger sizes (the 4x4 closed form of nalgebra is replaced by LU here: same value to
 * rounding). */
int orc_inverse(int n, const double* A, double* Ai) {
    if (n == 1) {
        if (A[0] == 0.0) return MPCB_INVERSE_FAIL;
        Ai[0] = 1.0 / A[0];
        return MPCB_OK;
    }
    if (n == 2) {
        const double m11 = A[0], m12 = A[1], m21 = A[2], m22 = A[3];
        const double det = m11 * m22 - m21 * m12;
        if (det == 0.0) return MPCB_INVERSE_FAIL;
        Ai[0] = m22 / det; Ai[1] = -m12 / det; Ai[2] = -m21 / det; Ai[3] = m11 / det;
        return MPCB_OK;
    }
    if (n == 3) {
        const double m11 = A[0], m12 = A[1], m13 = A[2], m21 = A[3], m22 = A[4], m23 = A[5], m31 = A[6], m32 = A[7],
                     m33 = A[8];
        const double minor_m12_m23 = m22 * m33 - m32 * m23;
        const double minor_m11_m23 = m21 * m33 - m31 * m23;
        const double minor_m11_m22 = m21 * m32 - m31 * m22;
        const double det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
        if (det == 0.0) return MPCB_INVERSE_FAIL;
        Ai[0] = minor_m12_m23 / det;
        Ai[1] = (m13 * m32 - m33 * m12) / det;
        Ai[2] = (m12 * m23 - m22 * m13) / det;
        Ai[3] = -minor_m11_m23 / det;
        Ai[4] = (m11 * m33 - m31 * m13) / det;
        Ai[5] = (m13 * m21 - m23 * m11) / det;
        Ai[6] = minor_m11_m22 / det;
        Ai[7] = (m12 * m31 - m32 * m11) / det;
        Ai[8] = (m11 * m22 - m21 * m12) / det;
        return MPCB_OK;
    }
    /* LU, partial pivoting, then solve for the identity columns */
    double lu[ORC_MAX_N * ORC_MAX_N];
    int perm[ORC_MAX_N];
    memcpy(lu, A, sizeof(double) * n * n);
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double best = fabs(lu[k * n + k]);
        for (int i = k + 1; i < n; ++i)
            if (fabs(lu[i * n + k]) > best) { best = fabs(lu[i * n + k]); piv = i; }
        if (best == 0.0) return MPCB_INVERSE_FAIL;
        if (piv != k) {
            for (int j = 0; j < n; ++j) { double t = lu[k * n + j]; lu[k * n + j] = lu[piv * n + j]; lu[piv * n + j] = t; }
            int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
        }
        const double d = lu[k * n + k];
        for (int i = k + 1; i < n; ++i) {
            lu[i * n + k] /= d;
            const double l = lu[i * n + k];
            for (int j = k + 1; j < n; ++j) lu[i * n + j] -= l * lu[k * n + j];
        }
    }
    for (int c = 0; c < n; ++c) {
        double y[ORC_MAX_N];
        for (int i = 0; i < n; ++i) {
            double s = (perm[i] == c) ? 1.0 : 0.0;
            for (int j = 0; j < i; ++j) s -= lu[i * n + j] * y[j];
            y[i] = s;
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = y[i];
            for (int j = i + 1; j < n; ++j) s -= lu[i * n + j] * Ai[j * n + c];
            Ai[i * n + c] = s / lu[i * n + i];
        }
    }
    return MPCB_OK;
}

int orc_ukf_sigma_points(int n, int sqrt_mode, int order, const double* x, const double* P, double* sig) {
    const int M = 2 * n + 1;
    const double C = ukf_c(n);
    double cp[ORC_MAX_N * ORC_MAX_N], L[ORC_MAX_N * ORC_MAX_N];
    for (int i = 0; i < n * n; ++i) cp[i] = C * P[i];
    if (sqrt_mode == MPCB_SQRT_CHOLESKY) {
        int st = orc_cholesky_lower(n, cp, L);
        if (st != MPCB_OK) return st;
    } else {
        orc_sym_eig_sqrt(n, cp, L);
    }
    for (int r = 0; r < n; ++r) {
        sig[r * M + 0] = x[r];
        for (int i = 0; i < n; ++i) {
            const int ip = (order == MPCB_ORDER_INTERLEAVED) ? 1 + 2 * i : 1 + i;
            const int im = (order == MPCB_ORDER_INTERLEAVED) ? 2 + 2 * i : 1 + n + i;
            sig[r * M + ip] = x[r] + L[r * n + i];
            sig[r * M + im] = x[r] - L[r * n + i];
        }
    }
    return MPCB_OK;
}

/* src/ukf.rs:96-110: mean = sigmas*wm (column-by-column axpy), P = sum_i wc_i*y_i*y_i^T + cov */
static void unscented_transform(int s, int M, const double* sig /*[s][M]*/, const double* wm, const double* wc,
                                const double* cov /*[s][s]*/, double* mean, double* P) {
    for (int r = 0; r < s; ++r) {
        double acc = sig[r * M + 0] * wm[0];
        for (int i = 1; i < M; ++i) acc += sig[r * M + i] * wm[i];
        mean[r] = acc;
    }
    for (int i = 0; i < s * s; ++i) P[i] = 0.0;
    for (int i = 0; i < M; ++i) {
        double y[ORC_MAX_N];
        for (int r = 0; r < s; ++r) y[r] = sig[r * M + i] - mean[r];
        for (int r = 0; r < s; ++r) {
            const double wy = wc[i] * y[r];
            for (int c = 0; c < s; ++c) P[r * s + c] += wy * y[c];
        }
    }
    for (int i = 0; i < s * s; ++i) P[i] = P[i] + cov[i];
}

int orc_ukf_predict(int model_id, const mpcb_model_params* p, int n, int sqrt_mode, int order, double* x, double* P,
                    const double* Q, double u, double dt, double* sigma_f) {
    const int M = 2 * n + 1;
    double wm[ORC_MAX_M], wc[ORC_MAX_M];
    orc_ukf_weights(n, wm, wc);
    int st = orc_ukf_sigma_points(n, sqrt_mode, order, x, P, sigma_f);
    if (st != MPCB_OK) return st;
    for (int i = 0; i < M; ++i) { /* :81-83 */
        double col[ORC_MAX_N], out[ORC_MAX_N];
        for (int r = 0; r < n; ++r) col[r] = sigma_f[r * M + i];
        orc_fx(model_id, p, col, u, dt, out);
        for (int r = 0; r < n; ++r) sigma_f[r * M + i] = out[r];
    }
    unscented_transform(n, M, sigma_f, wm, wc, Q, x, P);
    return MPCB_OK;
}

/* gen_r of examples/mppi4-ukf-commu.rs:228-236: the variance of every disabled sensor becomes 1e6 */
void orc_gen_r(int o, const double* R, uint32_t enable, double* R_out) {
    for (int i = 0; i < o * o; ++i) R_out[i] = R[i];
    for (int i = 0; i < o; ++i)
        if ((enable & (1u << i)) == 0) R_out[i * o + i] = 1e6;
}

int orc_ukf_update(int model_id, const mpcb_model_params* p, int n, int o, double* x, double* P, const double* R,
                   const double* z, const double* sigma_f) {
    return orc_ukf_update_masked(model_id, p, n, o, x, P, R, z, sigma_f, 0xffffffffu);
}

/* update with the per-packet sensor mask of examples/mppi4-ukf-commu.rs:279-293: the hx closure zeroes the
 * rows of disabled sensors (the caller pairs it with orc_gen_r) */
int orc_ukf_update_masked(int model_id, const mpcb_model_params* p, int n, int o, double* x, double* P, const double* R,
                          const double* z, const double* sigma_f, uint32_t enable) {
    const int M = 2 * n + 1;
    double wm[ORC_MAX_M], wc[ORC_MAX_M];
    orc_ukf_weights(n, wm, wc);
    double sh[ORC_MAX_N * ORC_MAX_M] = {0}; /* [o][M] */
    for (int i = 0; i < M; ++i) { /* :58-61 */
        double col[ORC_MAX_N], zz[ORC_MAX_N];
        for (int r = 0; r < n; ++r) col[r] = sigma_f[r * M + i];
        orc_hx(model_id, p, col, zz);
        for (int r = 0; r < o; ++r) sh[r * M + i] = (enable & (1u << r)) ? zz[r] : 0.0;
    }
    double zp[ORC_MAX_N], pz[ORC_MAX_N * ORC_MAX_N];
    unscented_transform(o, M, sh, wm, wc, R, zp, pz); /* :62 */
    double pxz[ORC_MAX_N * ORC_MAX_N]; /* [n][o] */
    for (int i = 0; i < n * o; ++i) pxz[i] = 0.0;
    for (int i = 0; i < M; ++i) { /* :63-68 */
        for (int r = 0; r < n; ++r) {
            const double wdx = wc[i] * (sigma_f[r * M + i] - x[r]);
            for (int c = 0; c < o; ++c) pxz[r * o + c] += wdx * (sh[c * M + i] - zp[c]);
        }
    }
    double pzi[ORC_MAX_N * ORC_MAX_N];
    int st = orc_inverse(o, pz, pzi); /* :69 */
    if (st != MPCB_OK) return st;
    double k[ORC_MAX_N * ORC_MAX_N]; /* [n][o] */
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < o; ++c) {
            double acc = pxz[r * o + 0] * pzi[0 * o + c];
            for (int j = 1; j < o; ++j) acc += pxz[r * o + j] * pzi[j * o + c];
            k[r * o + c] = acc;
        }
    /* :70 x += K (z - zp) */
    double innov[ORC_MAX_N];
    for (int c = 0; c < o; ++c) innov[c] = z[c] - zp[c];
    for (int r = 0; r < n; ++r) {
        double acc = k[r * o + 0] * innov[0];
        for (int j = 1; j < o; ++j) acc += k[r * o + j] * innov[j];
        x[r] += acc;
    }
    /* :71 P -= (K Pz) K^T */
    double kp[ORC_MAX_N * ORC_MAX_N]; /* [n][o] */
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < o; ++c) {
            double acc = k[r * o + 0] * pz[0 * o + c];
            for (int j = 1; j < o; ++j) acc += k[r * o + j] * pz[j * o + c];
            kp[r * o + c] = acc;
        }
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) {
            double acc = kp[r * o + 0] * k[c * o + 0];
            for (int j = 1; j < o; ++j) acc += kp[r * o + j] * k[c * o + j];
            P[r * n + c] -= acc;
        }
    /* :73 P = (P + P^T)/2 */
    for (int r = 0; r < n; ++r)
        for (int c = r; c < n; ++c) {
            const double a = (P[r * n + c] + P[c * n + r]) / 2.0;
            P[r * n + c] = a;
            P[c * n + r] = a;
        }
    return MPCB_OK;
}

int orc_ukf_step_batch(int model_id, const mpcb_model_params* p, int n, int o, int sqrt_mode, int order, int64_t B,
                       double* x, double* P, const double* Q, const double* R, const double* u, double u_scalar,
                       double dt, const double* z, int32_t* status, int threads) {
    return orc_ukf_step_batch_masked(model_id, p, n, o, sqrt_mode, order, B, x, P, Q, R, u, u_scalar, dt, z, status, threads,
                                     0xffffffffu);
}

int orc_ukf_step_batch_masked(int model_id, const mpcb_model_params* p, int n, int o, int sqrt_mode, int order, int64_t B,
                              double* x, double* P, const double* Q, const double* R, const double* u, double u_scalar,
                              double dt, const double* z, int32_t* status, int threads, uint32_t enable) {
    if (threads <= 0) threads = orc_max_threads();
    int any = MPCB_OK;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int64_t b = 0; b < B; ++b) {
        double sf[ORC_MAX_N * ORC_MAX_M];
        int st = orc_ukf_predict(model_id, p, n, sqrt_mode, order, x + b * n, P + b * n * n, Q, u ? u[b] : u_scalar, dt, sf);
        if (st == MPCB_OK) st = orc_ukf_update_masked(model_id, p, n, o, x + b * n, P + b * n * n, R, z + b * o, sf, enable);
        if (status) status[b] = st;
        if (st != MPCB_OK) {
#pragma omp critical
            if (any == MPCB_OK) any = st;
        }
    }
    return any;
}

/* ------------------------------------------------------------------ Gaussian ---- */
/* src/gaussian.rs:22-63 */
orc_gaussian orc_gaussian_add(orc_gaussian a, orc_gaussian b) {
    orc_gaussian r = {a.mean + b.mean, a.var + b.var};
    return r;
}
orc_gaussian orc_gaussian_sub(orc_gaussian a, orc_gaussian b) {
    orc_gaussian r = {a.mean - b.mean, a.var - b.var};
    return r;
}
orc_gaussian orc_gaussian_mul(orc_gaussian a, orc_gaussian b) {
    orc_gaussian r;
    r.mean = (a.var * b.mean + b.var * a.mean) / (a.var + b.var);
    r.var = (a.var * b.var) / (a.var + b.var);
    return r;
}
orc_gaussian orc_gaussian_scale(orc_gaussian a, double s) {
    orc_gaussian r = {a.mean * s, a.var * s};
    return r;
}
