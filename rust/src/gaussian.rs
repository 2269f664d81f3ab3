//! `mpc::gaussian::Gaussian` (reference src/gaussian.rs) — scalar host code, unchanged in meaning: it is not on the
//! GPU path. Source only.
#[derive(Debug, Clone, Copy, Default)]
pub struct Gaussian { pub mean: f64, pub var: f64 }

impl Gaussian { pub fn new(mean: f64, var: f64) -> Self { Self { mean, var } } }

impl core::ops::Add for Gaussian {
    type Output = Self;
    fn add(self, r: Self) -> Self { Self { mean: self.mean + r.mean, var: self.var + r.var } }
}
impl core::ops::Sub for Gaussian {
    type Output = Self;
    fn sub(self, r: Self) -> Self { Self { mean: self.mean - r.mean, var: self.var - r.var } }
}
impl core::ops::Mul for Gaussian {
    type Output = Self;
    fn mul(self, r: Self) -> Self {
        let s = self.var + r.var;
        Self { mean: (self.var * r.mean + r.var * self.mean) / s, var: self.var * r.var / s }
    }
}
impl core::ops::Mul<f64> for Gaussian {
    type Output = Self;
    fn mul(self, k: f64) -> Self { Self { mean: self.mean * k, var: self.var * k } }
}
