//! `mpc::ukf2::UnscentedKalmanFilter` (reference src/ukf2.rs, N = 6, O = 5) over the C ABI. Source only.
use crate::{ffi, DeviceModel};

const N: usize = 6;
const O: usize = 5;
type State = na::SVector<f64, N>;
type Cov<const D: usize> = na::SMatrix<f64, D, D>;

pub struct UnscentedKalmanFilter { h: *mut ffi::MpcbUkf, model: DeviceModel }
unsafe impl Send for UnscentedKalmanFilter {}

impl UnscentedKalmanFilter {
    /// src/ukf2.rs:30-42. nalgebra matrices are column-major; P, Q, R of a filter are symmetric, so passing them as the
    /// row-major arrays the ABI expects needs no transpose.
    pub fn new(x: State, p: Cov<N>, q: Cov<N>, r: Cov<O>) -> Self { Self::with_model(DeviceModel::Pen6, x, p, q, r) }

    pub fn with_model(model: DeviceModel, x: State, p: Cov<N>, q: Cov<N>, r: Cov<O>) -> Self {
        let mut cfg = unsafe { std::mem::zeroed::<ffi::MpcbUkfCfg>() };
        assert_eq!(unsafe { ffi::mpcb_ukf_default_cfg(model as i32, &mut cfg) }, 0);
        cfg.batch = 1;
        let mut h = std::ptr::null_mut();
        assert_eq!(unsafe { ffi::mpcb_ukf_create(&mut h, &cfg) }, 0, "mpcb_ukf_create: {}", ffi::last_error());
        assert_eq!(unsafe { ffi::mpcb_ukf_init(h, x.as_ptr(), p.as_ptr(), q.as_ptr(), r.as_ptr()) }, 0);
        Self { h, model }
    }

    /// src/ukf2.rs:44-52 — `fx` names the device model the filter was built with.
    pub fn predict(&mut self, u: f64, fx: DeviceModel) {
        assert_eq!(fx, self.model);
        assert_eq!(unsafe { ffi::mpcb_ukf_predict(self.h, std::ptr::null(), u, 0.0) }, 0);
        self.check();
    }

    /// src/ukf2.rs:54-74 — panics with "Inverse fail" like the reference's `.expect` (:69).
    pub fn update(&mut self, x_obs: &na::SVector<f64, O>, hx: DeviceModel) {
        assert_eq!(hx, self.model);
        assert_eq!(unsafe { ffi::mpcb_ukf_update(self.h, x_obs.as_ptr()) }, 0, "{}", ffi::last_error());
        self.check();
    }

    pub fn state(&self) -> State {
        let mut x = State::zeros();
        unsafe { ffi::mpcb_ukf_get_state(self.h, x.as_mut_ptr(), std::ptr::null_mut()) };
        x
    }

    pub fn covariance(&self) -> Cov<N> {
        let mut p = Cov::<N>::zeros();
        unsafe { ffi::mpcb_ukf_get_state(self.h, std::ptr::null_mut(), p.as_mut_ptr()) };
        p
    }

    pub fn set_q(&mut self, q: Cov<N>) { unsafe { ffi::mpcb_ukf_set_q(self.h, q.as_ptr()) }; }
    /// Called by examples/mppi4-ukf-commu.rs:280 and examples/mpc-ukf-commu.rs:334 but missing in the reference.
    pub fn set_r(&mut self, r: Cov<O>) { unsafe { ffi::mpcb_ukf_set_r(self.h, r.as_ptr()) }; }
    /// The `hx` closure of examples/mppi4-ukf-commu.rs:279-293 zeroes the rows of sensors whose enable bit is 0; a
    /// kernel cannot run that closure, so the mask is handed over instead (applies to the following `update`s).
    pub fn set_enable(&mut self, enable: u8) { unsafe { ffi::mpcb_ukf_set_enable(self.h, enable as u32) }; }

    fn check(&self) {
        let mut s = 0i32;
        match unsafe { ffi::mpcb_ukf_get_status(self.h, &mut s) } {
            0 => (),
            4 => panic!("Inverse fail"),
            5 => panic!("Cholesky fail"),
            _ => panic!("libmpc_b200: {}", ffi::last_error()),
        }
    }
}

impl Drop for UnscentedKalmanFilter {
    fn drop(&mut self) { unsafe { ffi::mpcb_ukf_destroy(self.h) } }
}
