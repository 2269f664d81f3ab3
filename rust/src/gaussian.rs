//! Host-side scalar type of the shim crate: `mpc::gaussian::Gaussian` of the reference (src/gaussian.rs, SURVEY.md 8a
//! row a12).  Nothing here touches the GPU — it exists so that a caller of the reference finds the same type and
//! operators after switching crates.  Source only (no Rust toolchain in the build image).
//!
//! Operator table (what each one means for a 1-D Kalman filter):
//!   a + b    prediction step: means and variances add
//!   a - b    means subtract, variances subtract as well (the reference's definition)
//!   a * b    measurement update: precision-weighted mean, harmonic combination of the variances
//!   a * k    scales the mean AND the variance by k (again the reference's definition, not k^2)
use core::ops;

/// Normal distribution N(mean, var).  `Default` is N(0, 0).
#[derive(Debug, Clone, Copy, Default, PartialEq)]
pub struct Gaussian {
    /// expected value
    pub mean: f64,
    /// variance (not the standard deviation)
    pub var: f64,
}

impl Gaussian {
    pub const fn new(mean: f64, var: f64) -> Gaussian {
        Gaussian { mean, var }
    }

    /// Product of two densities, renormalised: the fusion rule behind `*`, in the reference's operation order
    /// (src/gaussian.rs:44-52) so that the shim rounds like the crate it replaces.
    fn fuse(prior: Gaussian, meas: Gaussian) -> Gaussian {
        let total = prior.var + meas.var;
        let weighted = prior.var * meas.mean + meas.var * prior.mean;
        Gaussian::new(weighted / total, prior.var * meas.var / total)
    }
}

macro_rules! elementwise {
    ($trait:ident, $method:ident, $op:tt) => {
        impl ops::$trait for Gaussian {
            type Output = Gaussian;
            fn $method(self, other: Gaussian) -> Gaussian {
                Gaussian::new(self.mean $op other.mean, self.var $op other.var)
            }
        }
    };
}
elementwise!(Add, add, +);
elementwise!(Sub, sub, -);

impl ops::Mul<Gaussian> for Gaussian {
    type Output = Gaussian;
    fn mul(self, other: Gaussian) -> Gaussian {
        Gaussian::fuse(self, other)
    }
}

impl ops::Mul<f64> for Gaussian {
    type Output = Gaussian;
    fn mul(self, factor: f64) -> Gaussian {
        Gaussian::new(factor * self.mean, factor * self.var)
    }
}
