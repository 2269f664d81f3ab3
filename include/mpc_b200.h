/*
 * mpc_b200.h — C ABI of libmpc_b200.so: the B200 (sm_100a) implementation of the two
 * data-parallel hot paths of teruyamato0731/mpc-rs.
 *
 *   MPPI   Mppi<N,K,S>::new / compute                      reference src/mppi.rs:16-30, 33-92
 *   UKF    UnscentedKalmanFilter::new/predict/update/...   reference src/ukf.rs:30-94, src/ukf2.rs:30-98
 *          free-function Cholesky UKF                      reference examples/ukf-pen.rs:36-141
 *
 * Every entry point is plain `extern "C"`: POD structs, pointers and sizes only.  A Rust
 * `extern "C"` block / bindgen, ctypes, or cgo can bind this file as is (INTEGRATION.md shows
 * the stub).  Handles are opaque, own their device buffers and stream, are Send but not Sync
 * (one caller at a time per handle, like `&mut self` in the reference).  The library never keeps
 * a caller pointer past the call and never aborts: every failure is an mpcb_status.
 *
 * The reference passes its models as host fn pointers / closures (src/mppi.rs:9-10,
 * src/ukf.rs:44-46,54-56).  A kernel cannot call those, so the models of the reference's
 * examples are built in and selected by id + a parameter block.
 */
#ifndef MPC_B200_H
#define MPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPCB_ABI_VERSION 1

/* ------------------------------------------------------------------------------------------
 * Status codes.  The first three map 1:1 to the Err(&'static str) returns of Mppi::compute
 * (src/mppi.rs:69,77,88); INVERSE_FAIL / CHOLESKY_FAIL replace the reference's panics
 * (src/ukf.rs:69 "Inverse fail", examples/ukf-pen.rs:45 "Cholesky fail").
 * ---------------------------------------------------------------------------------------- */
typedef enum mpcb_status {
    MPCB_OK = 0,
    MPCB_NO_FINITE_COST = 1, /* "Cannot calculate max"  src/mppi.rs:69 */
    MPCB_SUM_ZERO = 2,       /* "sum is zero"           src/mppi.rs:77 */
    MPCB_U_INVALID = 3,      /* "u is invalid"          src/mppi.rs:88 */
    MPCB_INVERSE_FAIL = 4,   /* "Inverse fail"          src/ukf.rs:69  */
    MPCB_CHOLESKY_FAIL = 5,  /* "Cholesky fail"         examples/ukf-pen.rs:45 */
    MPCB_BAD_ARG = 6,
    MPCB_CUDA_ERROR = 7,
    MPCB_NCCL_ERROR = 8,
    MPCB_NOT_PREDICTED = 9,  /* update() before any predict(): the reference yields NaN (src/ukf.rs:32) */
    MPCB_PEER_TIMEOUT = 10,  /* multi-GPU exchange: a peer's partial row did not arrive within 20 s */
    MPCB_RTC_ERROR = 11      /* user model: NVRTC missing or the source does not compile (see mpcb_rtc_log) */
} mpcb_status;

/* Message for a status; for MPCB_NO_FINITE_COST/SUM_ZERO/U_INVALID/INVERSE_FAIL/CHOLESKY_FAIL it is
 * byte-identical to the reference's string. */
const char* mpcb_status_string(mpcb_status s);
/* Detail of the last CUDA/NCCL/argument error on the calling thread ("" if none). */
const char* mpcb_last_error_string(void);
int mpcb_abi_version(void);
/* Number of visible CUDA devices (<=0: none; the library has no CPU fallback). */
int mpcb_device_count(void);

/* ------------------------------------------------------------------------------------------
 * Built-in models.
 * ---------------------------------------------------------------------------------------- */
typedef enum mpcb_model {
    /* MPPI dynamics+cost, S = 4, state [x, x', th, th'] */
    MPCB_MODEL_L = 0,   /* linear cart-pendulum + clamped cost     examples/mppi4.rs:20-27,73-89 */
    MPCB_MODEL_NL = 1,  /* nonlinear pendulum, explicit Euler      examples/mppi4-non-liner.rs:20-27,73-94 */
    MPCB_MODEL_NL6 = 2, /* ddot + dynamics4, quadratic cost        examples/mppi4-non-liner-ukf.rs:33-35,126-148 */
    MPCB_MODEL_USER = 3, /* dynamics + cost supplied as CUDA source to mpcb_mppi_create_user                       */
    /* UKF process/measurement models */
    MPCB_MODEL_PEN_LIN = 16, /* n=4,o=2  fx/hx of examples/ukf-pen.rs:76-91 */
    MPCB_MODEL_PEN_NL = 17,  /* n=4,o=3  fx/hx of examples/ukf-pen2.rs:31-53 */
    MPCB_MODEL_PEN6 = 18,    /* n=6,o=5  fx/hx of examples/ukf-pen3.rs:35-63 (x[2].cos() as written) */
    MPCB_MODEL_NL6_UKF = 19, /* n=6,o=5  dynamics_short(.,.,dt,0)/hx of examples/mppi4-non-liner-ukf.rs:149-179 */
    MPCB_MODEL_USER_UKF = 20 /* fx + hx supplied as CUDA source to mpcb_ukf_create_user */
} mpcb_model;

/* Physical constants of the two-wheeled pendulum + cost weights.  mpcb_model_defaults() fills the
 * values each example ships with, evaluated in the example's own association order. */
typedef struct mpcb_model_params {
    double m1;   /* wheel mass            M1  */
    double r_w;  /* wheel radius          R_W */
    double m2;   /* pendulum mass         M2  */
    double l;    /* distance to c.o.g.    L   */
    double j1;   /* wheel inertia         J1  */
    double j2;   /* pendulum inertia      J2  */
    double g;    /* gravity               G   */
    double kt;   /* motor constant        KT  */
    double dt;   /* integration step      DT (MPPI: T/N; UKF: default dt of predict) */
    /* cost[0..3] state weights; L/NL additionally cost[4] |x0| clamp (2.0), cost[5] term2 clamp (5.0),
     * cost[6] x0 coupling in term2 (2.0), cost[7] x0 coupling in term3 (0.35), cost[8] its clamp (0.75)
     * (examples/mppi4.rs:20-27).  NL6 uses cost[0..3] only (examples/mppi4-non-liner-ukf.rs:21,33-35). */
    double cost[12];
} mpcb_model_params;

mpcb_status mpcb_model_defaults(int32_t model_id, mpcb_model_params* out);

typedef enum mpcb_precision {
    MPCB_F32 = 0, /* FP32 dynamics/cost/noise, FP64 cost accumulation + softmax + weighted mean */
    MPCB_F64 = 1, /* everything FP64, no FMA contraction: reference arithmetic */
    /* Everything FP64 in the folded form of the FP32 kernels: host-folded constants, one reciprocal per step instead of
     * the reference's divisions, sincos(), FMA contraction.  Differs from the reference order by FP64 rounding times
     * the model's error growth (model NL6 at its shipped DT: ~1e-11 on the controls); about half the FP64 instructions
     * of MPCB_F64.  Built-in models only; replay / dump noise is f64 like MPCB_F64. */
    MPCB_F64_FAST = 2
} mpcb_precision;

typedef enum mpcb_dtype { MPCB_DT_F32 = 0, MPCB_DT_F64 = 1 } mpcb_dtype;

/* ------------------------------------------------------------------------------------------
 * MPPI  (replaces mpc::mppi::Mppi<N,K,S>, src/mppi.rs:7-92)
 * ---------------------------------------------------------------------------------------- */
typedef struct mpcb_mppi mpcb_mppi;

typedef struct mpcb_mppi_cfg {
    int32_t model_id;    /* MPCB_MODEL_L / NL / NL6                                             */
    int32_t precision;   /* mpcb_precision                                                      */
    int32_t horizon;     /* N of Mppi<N,K,S> (1..512)                                           */
    int32_t state_dim;   /* S of Mppi<N,K,S>; the built-in models have S = 4, user models 1..8  */
    int64_t samples;     /* K of Mppi<N,K,S>: GLOBAL sample count per controller                */
    int32_t controllers; /* C independent controllers batched in one call (1 = the reference)   */
    int32_t device;      /* CUDA device ordinal                                                 */
    int32_t rank;        /* this handle computes global samples [rank*K/world, (rank+1)*K/world) */
    int32_t world_size;  /* 1 = single GPU                                                      */
    double lambda;       /* src/mppi.rs:11  */
    double std_dev;      /* src/mppi.rs:12 — a standard deviation (sigma), not a variance       */
    double limit_lo;     /* src/mppi.rs:8 limit.0 */
    double limit_hi;     /* src/mppi.rs:8 limit.1 */
    uint64_t seed;       /* Philox key for generate mode (the reference seeds from OS entropy, src/mppi.rs:41) */
    int32_t keep_costs;  /* 1: keep c_k (src/mppi.rs:61) of the last compute for mpcb_mppi_get_costs */
    int32_t reserved;
    mpcb_model_params model;
} mpcb_mppi_cfg;

/* Diagnostics of one compute (one per controller). */
typedef struct mpcb_mppi_info {
    int32_t status;    /* mpcb_status of this controller */
    int32_t reserved;
    int64_t argmax;    /* global index of the sample with the largest c_k = lowest cost (lowest index on ties); -1 if none finite */
    double max;        /* max finite c_k                        src/mppi.rs:65-69 */
    double sum;        /* sum_k exp((c_k - max)/lambda)         src/mppi.rs:74    */
    int64_t n_finite;  /* number of finite c_k                  src/mppi.rs:67    */
} mpcb_mppi_info;

mpcb_status mpcb_mppi_default_cfg(int32_t model_id, mpcb_mppi_cfg* out); /* the example's constants */

/* User-supplied model.  Mppi::new takes `dynamics: fn(&SVector<f64,S>, f64) -> SVector<f64,S>` and
 * `cost: fn(&SVector<f64,S>) -> f64` (src/mppi.rs:9-10,16-22); a kernel cannot call host functions, so the same two
 * functions are given as CUDA C++ source and the fused MPPI kernel is compiled around them at create time (NVRTC,
 * sm_100a; a few seconds).  `cuda_source` must define, at global scope (`__device__` may be omitted),
 *
 *     template <typename real> void dynamics(real (&x)[S], real u, const real* p);   // x <- f(x, u), in place
 *     template <typename real> real cost(const real (&x)[S], const real* p);         // stage cost of the NEW state
 *
 * with S = cfg->state_dim written as a literal (1..8: Mppi<N,K,S> is generic in S; the built-in models have S = 4)
 * (plain `float` / `double` overloads work too; `real` is float for MPCB_F32, double for MPCB_F64, where the source
 * is compiled without FMA contraction).  p[0..n_params) are the caller's constants (n_params <= MPCB_USER_PARAMS),
 * delivered through the kernel-parameter bank.  Helpers of the built-in models may be used: mpcb::sincos_r(a, &s, &c),
 * mpcb::fast_rcp(d), mpcb::clampm(v, lo, hi).  Everything else — noise, clamping of the control, control term,
 * softmax, weighted mean, errors, replay / dump / sharding — is the built-in path.  cfg->model_id and cfg->model are
 * ignored.  MPCB_RTC_ERROR if NVRTC is unavailable or the source does not compile. */
#define MPCB_USER_PARAMS 24
mpcb_status mpcb_mppi_create_user(mpcb_mppi** out, const mpcb_mppi_cfg* cfg, const char* cuda_source,
                                  const double* params, int32_t n_params);
/* Compile-only check of a user model with state dimension `state_dim` for `precision` (needs NVRTC, not a GPU). */
mpcb_status mpcb_mppi_check_user_source(const char* cuda_source, int32_t state_dim, int32_t precision);
/* NVRTC log (errors and warnings, lines of the user source as "user_model.cu(line)") of the calling thread's last
 * mpcb_mppi_create_user / mpcb_mppi_check_user_source; "" if none. */
const char* mpcb_rtc_log(void);
mpcb_status mpcb_mppi_create(mpcb_mppi** out, const mpcb_mppi_cfg* cfg); /* Mppi::new, src/mppi.rs:16 */
void mpcb_mppi_destroy(mpcb_mppi* h);

/* Mppi::compute (src/mppi.rs:33): host buffers x[C][S], u_in[C][H] -> u_out[C][H]; noise is generated
 * in-register (Philox4x32-10, key = (seed, call counter), counter = (global sample, t/4, controller)).
 * info may be NULL.  Returns the status of controller 0 when C == 1, else MPCB_OK unless the call
 * itself failed (per-controller status in info[c].status; a failed controller's u_out row is zeroed
 * like the caller-side fallback of examples/mppi4-non-liner-ukf.rs:80-86 when C > 1). */
mpcb_status mpcb_mppi_compute(mpcb_mppi* h, const double* x, const double* u_in, double* u_out,
                              mpcb_mppi_info* info);

/* Replay mode: the same computation on caller-supplied noise eps[C][K][H] (sample-major like the
 * reference's Vec<SVector<f64,N>>, src/mppi.rs:39-45), eps ~ N(0, std_dev^2) BEFORE the u_n shift and
 * clamp.  eps_dtype is mpcb_dtype; eps_on_device != 0 means eps is a device pointer on cfg.device.
 * With world_size > 1 eps still holds all K global samples; the handle reads its own shard. */
mpcb_status mpcb_mppi_compute_replay(mpcb_mppi* h, const double* x, const double* u_in,
                                     const void* eps, int32_t eps_dtype, int32_t eps_on_device,
                                     double* u_out, mpcb_mppi_info* info);

/* Generate mode that also returns the noise it drew (host buffer eps_out[C][K_local][H], mpcb_dtype
 * per the handle's precision: f32 for MPCB_F32, f64 for MPCB_F64) so the draw can be replayed. */
mpcb_status mpcb_mppi_compute_dump(mpcb_mppi* h, const double* x, const double* u_in, void* eps_out,
                                   double* u_out, mpcb_mppi_info* info);

/* c_k of the last compute (needs cfg.keep_costs): host buffer c[C][K_local]. */
mpcb_status mpcb_mppi_get_costs(mpcb_mppi* h, double* c_out);

/* Device-resident, asynchronous form (no host copies, no host sync): d_x[C][S], d_u_in[C][H],
 * d_u_out[C][H] are device pointers; work is enqueued on the handle's stream.  d_eps may be NULL
 * (generate).  Status lands in the handle; read it with mpcb_mppi_last_info after mpcb_mppi_sync. */
mpcb_status mpcb_mppi_compute_device(mpcb_mppi* h, const double* d_x, const double* d_u_in,
                                     const void* d_eps, int32_t eps_dtype, double* d_u_out);
/* d_u0[c] = d_u_out[c][0] on the handle's stream: the control every controller applies next (u_n[0],
 * examples/mppi4-non-liner-ukf.rs:231,270) — feeds mpcb_ukf_run_device's per-filter d_u without a host trip. */
mpcb_status mpcb_mppi_first_control_device(mpcb_mppi* h, const double* d_u_out, double* d_u0);
mpcb_status mpcb_mppi_sync(mpcb_mppi* h);
mpcb_status mpcb_mppi_last_info(mpcb_mppi* h, mpcb_mppi_info* info /*[C]*/);
void* mpcb_mppi_stream(mpcb_mppi* h);    /* cudaStream_t */
/* Global index of this handle's first controller (default 0): controllers sharded over GPUs draw the noise of their global
 * index, so a sharded batch reproduces the unsharded one. */
mpcb_status mpcb_mppi_set_controller_offset(mpcb_mppi* h, int64_t first_controller);
int64_t mpcb_mppi_launches(mpcb_mppi* h); /* kernels launched by this handle so far */
int64_t mpcb_mppi_local_samples(mpcb_mppi* h);

/* Multi-GPU: each rank reduces its sample shard to one partial row per controller
 * [max, sum_w, argmax, n_finite, sum_w*v[0..H)] (mpcb_mppi_partial_len doubles), the rows of all ranks are
 * exchanged (one small all-gather), and every rank combines them identically. */
int32_t mpcb_mppi_partial_len(mpcb_mppi* h); /* doubles per controller row */
/* d_partial[C][partial_len] device pointer; host x/u_in like mpcb_mppi_compute; d_eps NULL = generate */
mpcb_status mpcb_mppi_compute_partial(mpcb_mppi* h, const double* x, const double* u_in,
                                      const void* d_eps, int32_t eps_dtype, double* d_partial);
/* d_partials[G][C][partial_len] device pointer -> host u_out[C][H], info[C] */
mpcb_status mpcb_mppi_combine(mpcb_mppi* h, const double* d_partials, int32_t n_ranks, double* u_out,
                              mpcb_mppi_info* info);
/* In-library exchange over NCCL (NVLink/NVSwitch): after attach, mpcb_mppi_compute / _replay do
 * partial -> ncclAllGather -> combine on the handle's stream.  id is the 128-byte ncclUniqueId made by
 * mpcb_comm_unique_id on rank 0 and distributed by the caller (file, torch.distributed, MPI ...). */
mpcb_status mpcb_comm_unique_id(char id[128]);
mpcb_status mpcb_mppi_attach_comm(mpcb_mppi* h, const char id[128]);
/* Fused exchange over peer memory (NVLink / NVSwitch), no collective call: every rank owns a mailbox in its own
 * HBM; the kernel's final block stores this rank's partial row straight into every peer's mailbox, releases one
 * flag per peer, waits for the peers' rows and combines them — one kernel per control step.
 *   1. every rank: mpcb_mppi_peer_handle(h, mine)       (128 bytes: CUDA IPC handle of the mailbox + owner info)
 *   2. the caller gathers the handles of all ranks in rank order (file, torch.distributed, MPI ...)
 *   3. every rank: mpcb_mppi_attach_peers(h, all)       (all = world_size x 128 bytes)
 * Ranks may be separate processes (cudaIpcOpenMemHandle) or handles of one process (cudaDeviceEnablePeerAccess).
 * After attach, mpcb_mppi_compute / _replay / _device use this path; it takes precedence over an attached NCCL
 * communicator.  All ranks must make the same sequence of compute calls.  A rank whose peers never arrive gets
 * MPCB_PEER_TIMEOUT in info.status after 20 s instead of hanging. */
#define MPCB_PEER_HANDLE_BYTES 128
mpcb_status mpcb_mppi_peer_handle(mpcb_mppi* h, char out[MPCB_PEER_HANDLE_BYTES]);
mpcb_status mpcb_mppi_attach_peers(mpcb_mppi* h, const char* handles /* [world_size][MPCB_PEER_HANDLE_BYTES] */);

/* ------------------------------------------------------------------------------------------
 * UKF  (replaces mpc::ukf::UnscentedKalmanFilter n=4,o=3, mpc::ukf2::… n=6,o=5, and the free
 * functions of examples/ukf-pen.rs n=4,o=2), batched over B independent filters.  FP64 throughout
 * (the unscented weights are ±1e6, src/ukf.rs:24-28).  Device layout is structure-of-arrays
 * [component][B]; the host-facing calls take/return array-of-structures rows like the reference.
 * ---------------------------------------------------------------------------------------- */
typedef struct mpcb_ukf mpcb_ukf;

typedef enum mpcb_sqrt_mode {
    MPCB_SQRT_CHOLESKY = 0, /* lower Cholesky of C*P        examples/ukf-pen.rs:44-57 */
    MPCB_SQRT_EIG = 1       /* U*sqrt(S) of the SVD of C*P  src/ukf.rs:120-124 (symmetric PSD: cyclic Jacobi) */
} mpcb_sqrt_mode;

typedef enum mpcb_sigma_order {
    MPCB_ORDER_LIBRARY = 0,    /* [x, x+L0..x+Ln-1, x-L0..x-Ln-1]   src/ukf.rs:125-131 */
    MPCB_ORDER_INTERLEAVED = 1 /* [x, x+L0, x-L0, x+L1, x-L1, ...]  examples/ukf-pen.rs:46-56 */
} mpcb_sigma_order;

typedef struct mpcb_ukf_cfg {
    int32_t model_id;    /* MPCB_MODEL_PEN_LIN / PEN_NL / PEN6 / NL6_UKF */
    int32_t n;           /* state dim (4 or 6) — must match the model */
    int32_t o;           /* observation dim (2, 3 or 5) — must match the model */
    int32_t sqrt_mode;   /* mpcb_sqrt_mode */
    int32_t sigma_order; /* mpcb_sigma_order */
    int32_t device;
    int32_t exact;       /* 0 (default): FMA contraction, symmetric-half covariance sums, shared sigma weight
                          * factored out — rounding-level (~1e-10) away from the reference order, ~1.5x faster;
                          * 1: the reference's operation order without FMA (bit-exact vs an f64 restatement
                          * wherever libm agrees) */
    int32_t reserved;
    int64_t batch;       /* B filters on this handle (the caller shards B across GPUs; no exchange) */
    mpcb_model_params model;
} mpcb_ukf_cfg;

mpcb_status mpcb_ukf_default_cfg(int32_t model_id, mpcb_ukf_cfg* out);
/* Q[n][n], R[o][o], P0[n][n] the model's example ships with (row-major). */
mpcb_status mpcb_ukf_default_noise(int32_t model_id, double dt, double* Q, double* R, double* P0);
mpcb_status mpcb_ukf_create(mpcb_ukf** out, const mpcb_ukf_cfg* cfg);
/* User-supplied process and measurement models.  predict(u, fx) / update(&z, hx) take closures
 * `fx: Fn(&SVector<f64,N>, f64) -> SVector<f64,N>`, `hx: Fn(&SVector<f64,N>) -> SVector<f64,O>` (src/ukf.rs:44-46,54-56);
 * here the two functions are CUDA C++ source compiled into the batched UKF kernel at create time (NVRTC), with
 * cfg->n in 1..6 and cfg->o in 1..5 written as literals in the signatures:
 *
 *     void fx(double (&x)[N], double u, double dt, const double* p);        // x <- f(x, u), in place
 *     void hx(const double (&x)[N], double (&z)[O], const double* p);       // z <- h(x)
 *
 * dt is the dt of mpcb_ukf_predict / mpcb_ukf_step (cfg->model.dt when that is 0); p[0..n_params) the caller's
 * constants (n_params <= MPCB_USER_PARAMS).  cfg->model_id and the physical fields of cfg->model are ignored; sqrt_mode, sigma_order and exact
 * (no FMA contraction) apply as for the built-in models.  mpcb_ukf_default_cfg(MPCB_MODEL_USER_UKF) gives the
 * library-UKF defaults (eigen square root, library sigma order) with n = o = 0 for the caller to fill in. */
mpcb_status mpcb_ukf_create_user(mpcb_ukf** out, const mpcb_ukf_cfg* cfg, const char* cuda_source, const double* params,
                                 int32_t n_params);
mpcb_status mpcb_ukf_check_user_source(const char* cuda_source, int32_t n, int32_t o);
void mpcb_ukf_destroy(mpcb_ukf* h);

/* UnscentedKalmanFilter::new(x,p,q,r) (src/ukf.rs:30): same x[n], P[n][n] (row-major) for every filter */
mpcb_status mpcb_ukf_init(mpcb_ukf* h, const double* x, const double* P, const double* Q, const double* R);
/* per-filter state: host x[B][n], P[B][n][n] (either may be NULL to leave it unchanged) */
mpcb_status mpcb_ukf_set_state(mpcb_ukf* h, const double* x, const double* P);
/* state() / covariance() (src/ukf.rs:88-94) for all filters: host x[B][n], P[B][n][n]; either may be NULL */
mpcb_status mpcb_ukf_get_state(mpcb_ukf* h, double* x, double* P);
/* first/count window of the same, so a caller can sample a huge batch */
mpcb_status mpcb_ukf_get_state_range(mpcb_ukf* h, int64_t first, int64_t count, double* x, double* P);
mpcb_status mpcb_ukf_set_q(mpcb_ukf* h, const double* Q); /* src/ukf2.rs:96 */
mpcb_status mpcb_ukf_set_r(mpcb_ukf* h, const double* R); /* called by examples/mppi4-ukf-commu.rs:280, missing in the reference */
/* Per-packet sensor gating of examples/mppi4-ukf-commu.rs:228-236,279-293: bit i of `enable` cleared means sensor i
 * delivered nothing — the hx closure returns 0 for that row (mpcb_ukf_set_enable, applies to the following
 * update/step calls; default all ones) and gen_r inflates its variance to 1e6 (mpcb_ukf_gen_r builds that R from
 * the nominal R[o][o]; pass it to mpcb_ukf_set_r, as the example does). */
mpcb_status mpcb_ukf_set_enable(mpcb_ukf* h, uint32_t enable);
mpcb_status mpcb_ukf_gen_r(const mpcb_ukf* h, uint32_t enable, const double* R, double* R_out);
/* d_out[B][n_idx] (device, row per filter) <- state components idx[0..n_idx) of every filter, on the handle's
 * stream: e.g. idx = {0,1,3,4} turns the 6-state estimate into the 4-state MPPI input of
 * examples/mppi4-non-liner-ukf.rs:78 for mpcb_mppi_compute_device. */
mpcb_status mpcb_ukf_gather_state_device(mpcb_ukf* h, int32_t n_idx, const int32_t* idx, double* d_out);
/* predict(u, fx) (src/ukf.rs:44): u[B] per-filter controls, or NULL with u_scalar broadcast; dt <= 0 uses
 * cfg.model.dt (NL6_UKF's fx is dynamics_short(.,.,dt,0), examples/mppi4-non-liner-ukf.rs:278). */
mpcb_status mpcb_ukf_predict(mpcb_ukf* h, const double* u, double u_scalar, double dt);
/* update(&z, hx) (src/ukf.rs:54): host z[B][o] */
mpcb_status mpcb_ukf_update(mpcb_ukf* h, const double* z);
/* fused predict+update in one kernel (the sigma points never leave registers) */
mpcb_status mpcb_ukf_step(mpcb_ukf* h, const double* u, double u_scalar, double dt, const double* z);
/* Device-resident, asynchronous: T fused steps; d_z[T][o][B] (SoA), d_u[T][B] or NULL (u_scalar). */
mpcb_status mpcb_ukf_run_device(mpcb_ukf* h, int32_t steps, const double* d_u, double u_scalar, double dt,
                                const double* d_z);
mpcb_status mpcb_ukf_sync(mpcb_ukf* h);
/* per-filter mpcb_status after the last call(s) (sticky until mpcb_ukf_init/set_state): host s[B];
 * the return value is the first non-OK status found (MPCB_OK if none). */
mpcb_status mpcb_ukf_get_status(mpcb_ukf* h, int32_t* s);
void* mpcb_ukf_stream(mpcb_ukf* h);
int64_t mpcb_ukf_launches(mpcb_ukf* h);
/* raw SoA device pointers (x: [n][B], P: [n*n][B]) for zero-copy coupling with other device code.  After a fused step
 * (mpcb_ukf_step / mpcb_ukf_run_device) P is exactly symmetric and only its LOWER triangle (rows r >= columns c, entry
 * r*n + c) is current in device memory; mpcb_ukf_get_state mirrors it. */
double* mpcb_ukf_device_x(mpcb_ukf* h);
double* mpcb_ukf_device_p(mpcb_ukf* h);

/* ------------------------------------------------------------------------------------------
 * Device helpers so a host language without a CUDA binding can drive the *_device entry points.
 * ---------------------------------------------------------------------------------------- */
mpcb_status mpcb_device_alloc(int32_t device, uint64_t bytes, void** out);
mpcb_status mpcb_device_free(int32_t device, void* p);
mpcb_status mpcb_device_upload(int32_t device, void* dst, const void* src, uint64_t bytes);
mpcb_status mpcb_device_download(int32_t device, void* dst, const void* src, uint64_t bytes);

/* ---- batched closed loop (BASELINE config #4): examples/mppi4-non-liner-ukf.rs:38-103,224-288 ------------------------
 * C independent robots, each with its own MPPI controller (model NL6, the constants of :13-24) and its own UKF (model
 * NL6_UKF, :161-221), on a fixed tick instead of the reference's four wall-clock threads.  One tick = plant
 * (dynamics_short with the 2 N push of :236-241) + sensor (hx + R*N(0,1), :169-190) + UKF predict/update (:272-283) +
 * MPPI compute on the estimate's [x0, x1, x3, x4] (:55-87) — five launches on the device, no host round trip; a robot
 * whose MPPI step fails gets a zero control sequence like `Err(e) => zeros` (:81-86).  Controllers shard over GPUs
 * without any exchange: give every rank its own handle with controller_offset = its first global robot index (the
 * noise counters use the global index, so trajectories do not depend on the sharding). */
typedef struct mpcb_closed_loop mpcb_closed_loop;
typedef struct mpcb_closed_loop_cfg {
    int64_t controllers;       /* robots on this handle */
    int64_t samples;           /* K per controller (the example ships 500000; BASELINE config #4 uses 8192) */
    int64_t controller_offset; /* global index of this handle's first robot */
    double tick_dt;            /* fixed tick, seconds (0.01: the example's sensor period, :267-268) */
    uint64_t seed;
    int32_t use_estimate;      /* 1: MPPI is fed the UKF estimate; 0: the truth (DEBUG_UKF = true, :31,55-57) */
    int32_t precision;         /* MPCB_F32 / MPCB_F64 / MPCB_F64_FAST for the MPPI rollouts; anything else = the model's default */
    int32_t exact_ukf;         /* mpcb_ukf_cfg.exact */
    int32_t device;
} mpcb_closed_loop_cfg;
mpcb_status mpcb_closed_loop_default_cfg(mpcb_closed_loop_cfg* out); /* C = 4096, K = 8192, tick 0.01 s */
mpcb_status mpcb_closed_loop_create(mpcb_closed_loop** out, const mpcb_closed_loop_cfg* cfg);
void mpcb_closed_loop_destroy(mpcb_closed_loop* h);
/* truth and estimate of every robot, host x6[C][6] (init_ukf(&init_x), :40,161-167); P[C][6][6] or NULL to keep P0 */
mpcb_status mpcb_closed_loop_set_state(mpcb_closed_loop* h, const double* x6, const double* P);
mpcb_status mpcb_closed_loop_set_truth(mpcb_closed_loop* h, const double* x6);       /* the plant's state only */
mpcb_status mpcb_closed_loop_set_controls(mpcb_closed_loop* h, const double* u_seq); /* u_n[C][H] (and u_n[0]) */
/* n ticks, asynchronous (all launches are enqueued; read back with mpcb_closed_loop_get) */
mpcb_status mpcb_closed_loop_tick(mpcb_closed_loop* h, int32_t n_ticks);
/* one tick with the sensor readings z[C][5] and/or the MPPI noise eps[C][K][H] supplied by the caller (host pointers,
 * either may be NULL) — verification against a CPU restatement of the same schedule; synchronous */
mpcb_status mpcb_closed_loop_tick_replay(mpcb_closed_loop* h, const double* z, const void* eps, int32_t eps_dtype);
mpcb_status mpcb_closed_loop_sync(mpcb_closed_loop* h);
/* host copies after a sync; any pointer may be NULL: truth x6[C][6], estimate x_est[C][6], last readings z[C][5], applied
 * control u0[C], control sequences u_seq[C][H], per-controller mpcb_status of the last MPPI step */
mpcb_status mpcb_closed_loop_get(mpcb_closed_loop* h, double* x6, double* x_est, double* z, double* u0, double* u_seq,
                                 int32_t* mppi_status);
mpcb_mppi* mpcb_closed_loop_mppi(mpcb_closed_loop* h); /* the handles inside (owned by the loop) */
mpcb_ukf* mpcb_closed_loop_ukf(mpcb_closed_loop* h);
int64_t mpcb_closed_loop_ticks(mpcb_closed_loop* h);
int64_t mpcb_closed_loop_launches(mpcb_closed_loop* h);

#ifdef __cplusplus
}
#endif
#endif /* MPC_B200_H */
