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
