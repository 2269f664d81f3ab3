/*
 * mpc_oracle.h — CPU restatement of the mpc-rs MPPI / UKF hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library; the product (libmpc_b200.so) never links, loads or calls it.
 *
 * PARITY UNPINNED: the reference has no tests, asserts or golden vectors and its
 * Rust toolchain and arithmetic crates (nalgebra 0.33.0, rand 0.8.5, rand_distr 0.4.3,
 * rand_xoshiro 0.6.0, rayon 1.10.0 — Cargo.lock) are not in this image, so this oracle cannot be
 * checked against outputs of the reference itself.  It is pinned instead by closed-form known-answer
 * tests (tests/test_oracle_*.py): unscented weights, linear-model UKF == Kalman filter,
 * Gaussian product == scalar KF update, MPPI algebraic identities, and an independent numpy
 * re-derivation (tests/ref_numpy.py) written separately from the same reference lines.
 *
 * All arithmetic is IEEE f64 compiled with -ffp-contract=off (Rust never contracts to FMA), in the
 * association order of the cited reference lines.
 */
#ifndef MPC_ORACLE_H
#define MPC_ORACLE_H

#include <stdint.h>
#include "../include/mpc_b200.h" /* model ids, mpcb_model_params, status codes (POD only) */

#ifdef __cplusplus
extern "C" {
#endif

/* ---- models ---- */
int orc_model_defaults(int model_id, mpcb_model_params* p);
/* MPPI dynamics / cost, S = 4 (examples/mppi4.rs:82-89, mppi4-non-liner.rs:84-93, mppi4-non-liner-ukf.rs:126-148) */
void orc_dynamics(int model_id, const mpcb_model_params* p, const double* x, double u, double* out);
double orc_cost(int model_id, const mpcb_model_params* p, const double* x);
/* ddot / dynamics_short of examples/mppi4-non-liner-ukf.rs:126-159 (plant + UKF process model) */
void orc_ddot(const mpcb_model_params* p, const double* x4, double u, double f, double* ddx, double* ddth);
void orc_dynamics_short(const mpcb_model_params* p, const double* x6, double u, double dt, double f, double* out);
/* UKF fx / hx by model id; dt <= 0 uses p->dt */
void orc_fx(int model_id, const mpcb_model_params* p, const double* x, double u, double dt, double* out);
void orc_hx(int model_id, const mpcb_model_params* p, const double* x, double* z);
int orc_model_dims(int model_id, int* n, int* o);
void orc_gen_q(double dt, double* q36); /* examples/mppi4-non-liner-ukf.rs:192-221 */
int orc_ukf_default_noise(int model_id, double dt, double* Q, double* R, double* P0);

/* ---- MPPI (src/mppi.rs:33-92), replay form: eps[K][H] ~ N(0, std_dev^2) supplied by the caller ---- */
typedef struct orc_mppi_out {
    int32_t status;
    int64_t argmax;
    double max;
    double sum;
    int64_t n_finite;
} orc_mppi_out;

/* f64 reference arithmetic.  c_out[K] optional. Returns mpcb_status. */
int orc_mppi_compute(int model_id, const mpcb_model_params* p, int64_t K, int H, double lambda, double std_dev,
                     double lo, double hi, const double* x, const double* u_n, const double* eps,
                     double* u_out, double* c_out, orc_mppi_out* info);
/* f32 twin: dynamics/cost/noise in float with FMA-free float ops, cost/control-term/softmax in double —
 * bounds what the FP32 kernel can achieve; eps given as float[K][H]. */
int orc_mppi_compute_f32(int model_id, const mpcb_model_params* p, int64_t K, int H, double lambda, double std_dev,
                         double lo, double hi, const double* x, const double* u_n, const float* eps,
                         double* u_out, double* c_out, orc_mppi_out* info);

/* ---- CPU baseline: the reference's own structure (src/mppi.rs:38-84): materialise v[K][H], six passes,
 * one xoshiro256+ stream + ziggurat normal per worker, `threads` workers (rayon analogue). ---- */
int orc_mppi_compute_cpu(int model_id, const mpcb_model_params* p, int64_t K, int H, double lambda, double std_dev,
                         double lo, double hi, const double* x, const double* u_n, uint64_t seed, int threads,
                         double* u_out, orc_mppi_out* info);
int orc_max_threads(void);
/* fills out[n] with N(0,1) draws from the baseline's xoshiro256+/ziggurat sampler (for its own KS test) */
void orc_normal_fill(uint64_t seed, double* out, int64_t n);

/* ---- UKF ---- */
void orc_ukf_weights(int n, double* wm, double* wc);                    /* src/ukf.rs:112-118 */
/* sigma points [n][M] row-major (element (r,i) at r*M+i); returns status */
int orc_ukf_sigma_points(int n, int sqrt_mode, int order, const double* x, const double* P, double* sig);
int orc_cholesky_lower(int n, const double* A, double* L);             /* nalgebra cholesky().l() */
void orc_sym_eig_sqrt(int n, const double* A, double* Lout);           /* U*sqrt(S), src/ukf.rs:121-124 */
int orc_inverse(int n, const double* A, double* Ainv);                 /* try_inverse */
/* predict (src/ukf.rs:44-52): x[n], P[n][n] updated in place, sigma_f[n][M] out */
int orc_ukf_predict(int model_id, const mpcb_model_params* p, int n, int sqrt_mode, int order, double* x, double* P,
                    const double* Q, double u, double dt, double* sigma_f);
/* update (src/ukf.rs:54-74) */
int orc_ukf_update(int model_id, const mpcb_model_params* p, int n, int o, double* x, double* P, const double* R,
                   const double* z, const double* sigma_f);
int orc_ukf_update_masked(int model_id, const mpcb_model_params* p, int n, int o, double* x, double* P, const double* R,
                          const double* z, const double* sigma_f, uint32_t enable); /* examples/mppi4-ukf-commu.rs:279-293 */
void orc_gen_r(int o, const double* R, uint32_t enable, double* R_out);            /* examples/mppi4-ukf-commu.rs:228-236 */
int orc_ukf_step_batch_masked(int model_id, const mpcb_model_params* p, int n, int o, int sqrt_mode, int order, int64_t B,
                              double* x, double* P, const double* Q, const double* R, const double* u, double u_scalar,
                              double dt, const double* z, int32_t* status, int threads, uint32_t enable);
/* batched predict+update over B filters, AoS x[B][n], P[B][n][n], z[B][o], u[B] or NULL; `threads` workers */
int orc_ukf_step_batch(int model_id, const mpcb_model_params* p, int n, int o, int sqrt_mode, int order, int64_t B,
                       double* x, double* P, const double* Q, const double* R, const double* u, double u_scalar,
                       double dt, const double* z, int32_t* status, int threads);

/* ---- Gaussian (src/gaussian.rs:22-63) ---- */
typedef struct orc_gaussian { double mean, var; } orc_gaussian;
orc_gaussian orc_gaussian_add(orc_gaussian a, orc_gaussian b);
orc_gaussian orc_gaussian_sub(orc_gaussian a, orc_gaussian b);
orc_gaussian orc_gaussian_mul(orc_gaussian a, orc_gaussian b);
orc_gaussian orc_gaussian_scale(orc_gaussian a, double s);

#ifdef __cplusplus
}
#endif
#endif
