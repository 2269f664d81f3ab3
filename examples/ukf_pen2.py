#!/usr/bin/env python
"""examples/ukf-pen2.rs (mpc::ukf, n = 4, o = 3) and examples/ukf-pen3.rs (mpc::ukf2, n = 6, o = 5, --six) on B200:
the library UKF with the SVD square root tracking the nonlinear pendulum for 100 open-loop steps, written against the
reference's method names (new / predict / update / state / covariance) and printing the reference's line.

    python examples/ukf_pen2.py [--six] [--steps 100]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_rs_b200 import UnscentedKalmanFilter, models, ukf  # noqa: E402
from mpc_rs_b200.plants import PlantPen6, PlantPenNL  # noqa: E402


def _fmt(v):
    return "(" + ",".join(f"{a:7.2f}" for a in v) + ")"


def run(six=False, steps=100, quiet=False, seed=0):
    model = models.PEN6 if six else models.PEN_NL
    plant = PlantPen6() if six else PlantPenNL()
    n = 6 if six else 4
    shown = [0, 1, 3, 4] if six else [0, 1, 2, 3]          # ukf-pen3.rs:101-104 / ukf-pen2.rs:89-92
    n_obs_shown = 5 if six else 2                          # ukf-pen3.rs:105-108 / ukf-pen2.rs:93
    rng = np.random.default_rng(seed)
    Q, R, P0 = ukf.default_noise(model)                    # ukf-pen2.rs:18-28,71-76 / ukf-pen3.rs:18-32,83-90
    x_act = np.zeros(n)
    f = UnscentedKalmanFilter.new(np.zeros(n), P0, Q, R, fx=model)
    u = 0.1
    for i in range(steps):
        x_act = plant.fx(x_act, u)
        f.predict(u, model)
        x_obs = plant.sensor(x_act, rng)
        f.update(x_obs, model)
        x_est, p = f.state(), f.covariance()
        if not quiet:
            print(f"t: {i * plant.dt:4.2f} x_act: {_fmt(x_act[shown])} x_obs: {_fmt(x_obs[:n_obs_shown])} "
                  f"x_est: {_fmt(x_est[shown])} p: {_fmt(np.diag(p))}")
    f.close()
    return x_act, x_est, p


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--six", action="store_true", help="examples/ukf-pen3.rs (mpc::ukf2, n = 6, o = 5)")
    ap.add_argument("--steps", type=int, default=100)
    a = ap.parse_args()
    run(a.six, a.steps)
