//! Hand-written declarations of include/mpc_b200.h (what `bindgen` would emit). Source only.
#![allow(non_camel_case_types)]
#[repr(C)] #[derive(Clone, Copy)]
pub struct MpcbModelParams { pub m1: f64, pub r_w: f64, pub m2: f64, pub l: f64, pub j1: f64, pub j2: f64,
    pub g: f64, pub kt: f64, pub dt: f64, pub cost: [f64; 12] }
#[repr(C)] #[derive(Clone, Copy)]
pub struct MpcbMppiCfg { pub model_id: i32, pub precision: i32, pub horizon: i32, pub state_dim: i32,
    pub samples: i64, pub controllers: i32, pub device: i32, pub rank: i32, pub world_size: i32, pub lambda: f64,
    pub std_dev: f64, pub limit_lo: f64, pub limit_hi: f64, pub seed: u64, pub keep_costs: i32, pub reserved: i32,
    pub model: MpcbModelParams }
#[repr(C)] #[derive(Clone, Copy)]
pub struct MpcbMppiInfo { pub status: i32, pub reserved: i32, pub argmax: i64, pub max: f64, pub sum: f64,
    pub n_finite: i64 }
#[repr(C)] #[derive(Clone, Copy)]
pub struct MpcbUkfCfg { pub model_id: i32, pub n: i32, pub o: i32, pub sqrt_mode: i32, pub sigma_order: i32,
    pub device: i32, pub exact: i32, pub reserved: i32, pub batch: i64, pub model: MpcbModelParams }
pub enum MpcbMppi {}
pub enum MpcbUkf {}
pub const MPCB_MODEL_USER_UKF: i32 = 20; // fx + hx supplied as CUDA source (mpcb_ukf_create_user)
pub const MPCB_MODEL_USER: i32 = 3; // dynamics + cost supplied as CUDA source (mpcb_mppi_create_user)
// mpcb_precision (include/mpc_b200.h): mpcb_mppi_default_cfg picks F32 for models L / NL and F64_FAST for model NL6
pub const MPCB_F32: i32 = 0; // FP32 rollouts, FP64 accumulation: 1e-5 from the f64 reference on L / NL
pub const MPCB_F64: i32 = 1; // the reference's own operation order, no FMA contraction
pub const MPCB_F64_FAST: i32 = 2; // FP64 in the folded form: 1e-9 from the reference at about half the cost of MPCB_F64

extern "C" {
    pub fn mpcb_last_error_string() -> *const std::os::raw::c_char;
    pub fn mpcb_mppi_default_cfg(model_id: i32, out: *mut MpcbMppiCfg) -> i32;
    pub fn mpcb_mppi_create(out: *mut *mut MpcbMppi, cfg: *const MpcbMppiCfg) -> i32;
    pub fn mpcb_mppi_create_user(out: *mut *mut MpcbMppi, cfg: *const MpcbMppiCfg, cuda_source: *const std::os::raw::c_char,
                                 params: *const f64, n_params: i32) -> i32;
    pub fn mpcb_mppi_check_user_source(cuda_source: *const std::os::raw::c_char, state_dim: i32, precision: i32) -> i32;
    pub fn mpcb_rtc_log() -> *const std::os::raw::c_char;
    pub fn mpcb_ukf_create_user(out: *mut *mut MpcbUkf, cfg: *const MpcbUkfCfg, cuda_source: *const std::os::raw::c_char,
                                params: *const f64, n_params: i32) -> i32;
    pub fn mpcb_mppi_destroy(h: *mut MpcbMppi);
    pub fn mpcb_mppi_compute(h: *mut MpcbMppi, x: *const f64, u_in: *const f64, u_out: *mut f64, info: *mut MpcbMppiInfo) -> i32;
    pub fn mpcb_ukf_default_cfg(model_id: i32, out: *mut MpcbUkfCfg) -> i32;
    pub fn mpcb_ukf_create(out: *mut *mut MpcbUkf, cfg: *const MpcbUkfCfg) -> i32;
    pub fn mpcb_ukf_destroy(h: *mut MpcbUkf);
    pub fn mpcb_ukf_init(h: *mut MpcbUkf, x: *const f64, p: *const f64, q: *const f64, r: *const f64) -> i32;
    pub fn mpcb_ukf_get_state(h: *mut MpcbUkf, x: *mut f64, p: *mut f64) -> i32;
    pub fn mpcb_ukf_set_q(h: *mut MpcbUkf, q: *const f64) -> i32;
    pub fn mpcb_ukf_set_r(h: *mut MpcbUkf, r: *const f64) -> i32;
    pub fn mpcb_ukf_predict(h: *mut MpcbUkf, u: *const f64, u_scalar: f64, dt: f64) -> i32;
    pub fn mpcb_ukf_update(h: *mut MpcbUkf, z: *const f64) -> i32;
    pub fn mpcb_ukf_get_status(h: *mut MpcbUkf, s: *mut i32) -> i32;
    // per-packet sensor gating (examples/mppi4-ukf-commu.rs:228-236,279-293)
    pub fn mpcb_ukf_set_enable(h: *mut MpcbUkf, enable: u32) -> i32;
    pub fn mpcb_ukf_gen_r(h: *const MpcbUkf, enable: u32, r: *const f64, r_out: *mut f64) -> i32;
    // multi-GPU: fused exchange over peer memory (128-byte handles gathered by the application) or NCCL
    pub fn mpcb_mppi_peer_handle(h: *mut MpcbMppi, out: *mut std::os::raw::c_char) -> i32;
    pub fn mpcb_mppi_attach_peers(h: *mut MpcbMppi, handles: *const std::os::raw::c_char) -> i32;
    pub fn mpcb_comm_unique_id(id: *mut std::os::raw::c_char) -> i32;
    pub fn mpcb_mppi_attach_comm(h: *mut MpcbMppi, id: *const std::os::raw::c_char) -> i32;
}
pub fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(mpcb_last_error_string()).to_string_lossy().into_owned() }
}
