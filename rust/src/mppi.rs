//! `mpc::mppi::Mppi<N,K,S>` (reference src/mppi.rs:7-92) over the C ABI. Source only.
use crate::{ffi, DeviceModel};

pub struct Mppi<const N: usize, const K: usize, const S: usize> { h: *mut ffi::MpcbMppi }
unsafe impl<const N: usize, const K: usize, const S: usize> Send for Mppi<N, K, S> {}

impl<const N: usize, const K: usize, const S: usize> Mppi<N, K, S> {
    /// Argument order of the reference's `new` (src/mppi.rs:16-22); `dynamics`/`cost` name a built-in device model.
    pub fn new(dynamics: DeviceModel, cost: DeviceModel, lambda: f64, std_dev: f64, limit: (f64, f64)) -> Self {
        assert_eq!(dynamics, cost, "dynamics and cost must name the same device model");
        let mut cfg = unsafe { std::mem::zeroed::<ffi::MpcbMppiCfg>() };
        assert_eq!(unsafe { ffi::mpcb_mppi_default_cfg(dynamics as i32, &mut cfg) }, 0);
        cfg.horizon = N as i32;
        cfg.samples = K as i64;
        cfg.state_dim = S as i32;
        cfg.lambda = lambda;
        cfg.std_dev = std_dev;
        cfg.limit_lo = limit.0;
        cfg.limit_hi = limit.1;
        let mut h = std::ptr::null_mut();
        let st = unsafe { ffi::mpcb_mppi_create(&mut h, &cfg) };
        assert_eq!(st, 0, "mpcb_mppi_create: {}", ffi::last_error());
        Self { h }
    }

    /// `new` with the caller's own `dynamics` and `cost` (src/mppi.rs:9-10), given as CUDA C++ source:
    ///     template <typename real> void dynamics(real (&x)[4], real u, const real* p);
    ///     template <typename real> real cost(const real (&x)[4], const real* p);
    /// compiled into the fused kernel at construction (mpcb_mppi_create_user); Err carries the compiler log.
    pub fn new_cuda(source: &str, params: &[f64], lambda: f64, std_dev: f64, limit: (f64, f64)) -> Result<Self, String> {
        let mut cfg = unsafe { std::mem::zeroed::<ffi::MpcbMppiCfg>() };
        assert_eq!(unsafe { ffi::mpcb_mppi_default_cfg(ffi::MPCB_MODEL_USER, &mut cfg) }, 0);
        cfg.horizon = N as i32;
        cfg.samples = K as i64;
        cfg.state_dim = S as i32;
        cfg.lambda = lambda;
        cfg.std_dev = std_dev;
        cfg.limit_lo = limit.0;
        cfg.limit_hi = limit.1;
        let src = std::ffi::CString::new(source).map_err(|e| e.to_string())?;
        let mut h = std::ptr::null_mut();
        match unsafe { ffi::mpcb_mppi_create_user(&mut h, &cfg, src.as_ptr(), params.as_ptr(), params.len() as i32) } {
            0 => Ok(Self { h }),
            11 => Err(unsafe { std::ffi::CStr::from_ptr(ffi::mpcb_rtc_log()) }.to_string_lossy().into_owned()),
            _ => Err(ffi::last_error()),
        }
    }

    /// src/mppi.rs:33-92; the three Err strings are the reference's (:69, :77, :88).
    pub fn compute(&mut self, x: &na::SVector<f64, S>, u_n: &na::SVector<f64, N>) -> Result<na::SVector<f64, N>, &'static str> {
        let mut out = na::SVector::<f64, N>::zeros();
        let st = unsafe { ffi::mpcb_mppi_compute(self.h, x.as_ptr(), u_n.as_ptr(), out.as_mut_ptr(), std::ptr::null_mut()) };
        match st {
            0 => Ok(out),
            1 => Err("Cannot calculate max"),
            2 => Err("sum is zero"),
            3 => Err("u is invalid"),
            _ => panic!("libmpc_b200: {}", ffi::last_error()),
        }
    }
}

impl<const N: usize, const K: usize, const S: usize> Drop for Mppi<N, K, S> {
    fn drop(&mut self) { unsafe { ffi::mpcb_mppi_destroy(self.h) } }
}
