#!/usr/bin/env python
"""Mppi::new(dynamics, cost, ...) with the caller's OWN model (src/mppi.rs:9-10,16-22): a cart that has to park at
x = 1 m while keeping a hanging load (pendulum pointing down, small-angle damped) quiet — a model the library does not
ship.  The two functions are CUDA C++ source compiled into the fused MPPI kernel at construction
(mpcb_mppi_create_user); the plant on the host is the same formula in numpy.

    python examples/mppi4_user_model.py [--samples 65536] [--seconds 6]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_rs_b200 import Mppi, user_model  # noqa: E402

SOURCE = r"""
// state [x, x', phi, phi'] (phi = swing angle of the load), control = cart acceleration; p = dt, g/l, damping, target
template <typename real>
void dynamics(real (&x)[4], real u, const real* p) {
    const real dt = p[0];
    real s, c;
    mpcb::sincos_r(x[2], &s, &c);
    const real acc_phi = -p[1] * s - u * c * p[1] / (real)9.81 - p[2] * x[3];
    x[0] += x[1] * dt;
    x[1] += u * dt;
    x[2] += x[3] * dt;
    x[3] += acc_phi * dt;
}
template <typename real>
real cost(const real (&x)[4], const real* p) {
    const real e = x[0] - p[3];
    return (real)4.0 * (e * e) + (real)0.5 * (x[1] * x[1]) + (real)6.0 * (x[2] * x[2]) + (real)0.4 * (x[3] * x[3]);
}
"""
N, DT = 20, 0.05
PARAMS = [DT, 9.81 / 0.8, 0.05, 1.0]


def plant(x, u):
    dt, gl, damp, _ = PARAMS
    acc = -gl * np.sin(x[2]) - u * np.cos(x[2]) * gl / 9.81 - damp * x[3]
    return np.array([x[0] + x[1] * dt, x[1] + u * dt, x[2] + x[3] * dt, x[3] + acc * dt])


def run(samples=65536, seconds=6.0, quiet=False, seed=1):
    mppi = Mppi.new(user_model(SOURCE, PARAMS), user_model(SOURCE, PARAMS), 0.3, 1.0, (-3.0, 3.0), N=N, K=samples, seed=seed)
    x, u_n = np.zeros(4), np.zeros(N)
    t, rows = 0.0, []
    while t < seconds:
        u_n = mppi.compute(x, u_n)  # Err(..) of the reference raises MppiError here
        x = plant(x, u_n[0])
        rows.append((t, u_n[0], *x))
        if not quiet:
            print(f"t: {t:.2f}, u: {u_n[0]:6.2f}, x: [{x[0]:6.2f}, {x[1]:5.2f}, {x[2]:5.2f}, {x[3]:5.2f}]")
        t += DT
    mppi.close()
    return np.array(rows)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=65536)
    ap.add_argument("--seconds", type=float, default=6.0)
    a = ap.parse_args()
    run(a.samples, a.seconds)
