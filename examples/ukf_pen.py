#!/usr/bin/env python
"""examples/ukf-pen.rs on B200: the free-function Cholesky UKF (n = 4, o = 2) tracking the linear pendulum for 100
steps, printing x_act / x_obs / x_est / diag(P) per step like the reference — optionally for a whole batch of
independent filters (BASELINE config #3 is --batch 1048576).

    python examples/ukf_pen.py [--batch B] [--steps 100]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_rs_b200 import BatchedUkf, models, ukf  # noqa: E402
from mpc_rs_b200.plants import PlantPenLin  # noqa: E402


def run(batch=1, steps=100, quiet=False, seed=0):
    plant = PlantPenLin()
    rng = np.random.default_rng(seed)
    Q, R, P0 = ukf.default_noise(models.PEN_LIN)  # examples/ukf-pen.rs:17-26,148-153
    x_act = np.zeros((batch, 4))
    u = 0.0015  # :155
    with BatchedUkf(models.PEN_LIN, batch) as f:  # Cholesky square root, interleaved sigma order (:44-57)
        f.init(np.zeros(4), P0, Q, R)
        for _ in range(steps):
            x_act = plant.fx(x_act, u)
            x_obs = plant.sensor(x_act, rng)
            f.step(u, x_obs)  # predict(&mut x_est, u, &mut p) + update(..)  (:156-159), fused
            if not quiet:
                x_est, p = f.get_state(0, 1)
                a, o, e, d = x_act[0], x_obs[0], x_est[0], np.diag(p[0])
                print(f"x_act: ({a[0]:7.2f},{a[1]:7.2f},{a[2]:7.2f},{a[3]:7.2f}) x_obs: ({o[0]:7.2f},{o[1]:7.2f}) "
                      f"x_est: ({e[0]:7.2f},{e[1]:7.2f},{e[2]:7.2f},{e[3]:7.2f}) p: ({d[0]:7.2f},{d[1]:7.2f},{d[2]:7.2f},{d[3]:7.2f})")
        x_est, p = f.get_state()
    return x_act, x_est, p


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    a = ap.parse_args()
    run(a.batch, a.steps)
