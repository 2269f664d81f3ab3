"""mpc::gaussian::Gaussian (reference src/gaussian.rs:1-63): 1-D Gaussian algebra.  Scalar host code, like the
reference (16-byte POD, 8 flops per op) — it is not on the GPU path."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Gaussian:
    mean: float = 0.0  # Default::default() is (0, 0), src/gaussian.rs:13-20
    var: float = 0.0

    @classmethod
    def new(cls, mean: float, var: float) -> "Gaussian":
        return cls(float(mean), float(var))

    def __add__(self, rhs: "Gaussian") -> "Gaussian":  # :22-31
        return Gaussian(self.mean + rhs.mean, self.var + rhs.var)

    def __sub__(self, rhs: "Gaussian") -> "Gaussian":  # :33-42 (the variance subtracts)
        return Gaussian(self.mean - rhs.mean, self.var - rhs.var)

    def __mul__(self, rhs):
        if isinstance(rhs, Gaussian):  # :44-52 product of Gaussians = scalar Kalman update
            mean = (self.var * rhs.mean + rhs.var * self.mean) / (self.var + rhs.var)
            var = (self.var * rhs.var) / (self.var + rhs.var)
            return Gaussian(mean, var)
        return Gaussian(self.mean * rhs, self.var * rhs)  # :54-63 scales mean AND var by the scalar
