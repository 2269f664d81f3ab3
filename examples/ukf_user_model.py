#!/usr/bin/env python
"""predict(u, fx) / update(&z, hx) with the caller's OWN closures (src/ukf.rs:44-46,54-56): a damped pendulum with an
unknown constant torque bias, n = 3 states [theta, omega, bias], o = 2 sensors (a sine-shaped angle pickup and a rate
gyro) — dimensions the reference's two filter types (4,3) and (6,5) do not cover.  fx / hx are CUDA C++ source compiled
into the batched UKF kernel at construction (mpcb_ukf_create_user); the plant on the host is the same formula in numpy.

    python examples/ukf_user_model.py [--batch B] [--steps 400]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_rs_b200 import BatchedUkf, user_ukf_model  # noqa: E402

SOURCE = r"""
// p = g/l, damping, pickup gain
void fx(double (&x)[3], double u, double dt, const double* p) {
    const double th = x[0], om = x[1];
    x[0] = th + om * dt;
    x[1] = om + (u + x[2] - p[0] * sin(th) - p[1] * om) * dt;
}
void hx(const double (&x)[3], double (&z)[2], const double* p) {
    z[0] = p[2] * sin(x[0]);
    z[1] = x[1];
}
"""
PARAMS = [9.81 / 0.6, 0.15, 2.0]
DT = 0.01
Q = np.diag([1e-6, 1e-4, 1e-6])
R = np.diag([0.02 ** 2, 0.05 ** 2])
P0 = np.diag([0.5, 1.0, 4.0])


def plant(x, u):
    gl, damp, _ = PARAMS
    th, om, b = x[..., 0], x[..., 1], x[..., 2]
    return np.stack([th + om * DT, om + (u + b - gl * np.sin(th) - damp * om) * DT, b], axis=-1)


def run(batch=1, steps=400, quiet=False, seed=0):
    rng = np.random.default_rng(seed)
    x_act = np.tile(np.array([0.6, 0.0, 1.5]), (batch, 1)) + rng.normal(0, 0.05, (batch, 3))  # true bias 1.5, unknown to the filter
    with BatchedUkf(user_ukf_model(SOURCE, 3, 2, PARAMS), batch, dt=DT) as f:
        f.init(np.array([0.5, 0.0, 0.0]), P0, Q, R)
        for i in range(steps):
            u = 0.8 * np.sin(0.01 * i * 2 * np.pi)
            x_act = plant(x_act, u)
            z = np.stack([PARAMS[2] * np.sin(x_act[:, 0]), x_act[:, 1]], axis=-1) + rng.normal(0, 1.0, (batch, 2)) * np.sqrt(np.diag(R))
            f.step(u, z, DT)
            if not quiet and i % 40 == 39:
                xe, p = f.get_state(0, 1)
                print(f"t: {i * DT:4.2f} x_act: ({x_act[0, 0]:6.3f},{x_act[0, 1]:6.3f},{x_act[0, 2]:6.3f}) "
                      f"x_est: ({xe[0, 0]:6.3f},{xe[0, 1]:6.3f},{xe[0, 2]:6.3f}) p: ({p[0, 0, 0]:8.1e},{p[0, 1, 1]:8.1e},{p[0, 2, 2]:8.1e})")
        x_est, p = f.get_state()
    return x_act, x_est, p


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    a = ap.parse_args()
    run(a.batch, a.steps)
