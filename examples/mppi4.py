#!/usr/bin/env python
"""examples/mppi4.rs on B200: the linear cart-pendulum MPPI loop as shipped (K = 800 000, N = 8, 10 s, stop at
|theta| > 60 deg), printing the same line per step and writing logs/mppi/mppi.csv in the reference's format.

    python examples/mppi4.py [--nonlinear] [--samples K] [--seconds 10] [--csv logs/mppi/mppi.csv]
--nonlinear runs examples/mppi4-non-liner.rs instead (same constants, nonlinear plant and rollouts)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_rs_b200 import Mppi, models  # noqa: E402
from mpc_rs_b200.csvlog import MppiLog  # noqa: E402
from mpc_rs_b200.plants import PlantL, PlantNL  # noqa: E402

T, N, LAMBDA, R, LIMIT = 0.8, 8, 0.5, 3.0, (-20.0, 20.0)  # examples/mppi4.rs:8-18
DT = T / N


def run(nonlinear=False, samples=800_000, seconds=10.0, csv="logs/mppi/mppi.csv", quiet=False, seed=None):
    model = models.NL if nonlinear else models.L
    plant = (PlantNL if nonlinear else PlantL)(DT)
    x = np.array([0.5, 0.0, 0.1, 0.0])  # :29
    u_n = np.zeros(N)
    mppi = Mppi.new(model, model, LAMBDA, R, LIMIT, N=N, K=samples, seed=seed)
    rows = []
    now = time.perf_counter()
    t = 0.0
    with MppiLog(csv) as log:
        while t < seconds:
            u_n = mppi.compute(x, u_n)
            x = plant.step(x, u_n[0])
            if not quiet:
                print(f"t: {t:.2f}, u: {u_n[0]:6.2f}, x: [{x[0]:6.2f}, {x[1]:5.2f}, {x[2]:5.2f}, {x[3]:5.2f}]")
            if abs(x[2]) > np.radians(60.0):
                print("x[2] is over 60 degrees")
                break
            log.write(t, u_n[0], x)
            rows.append((t, u_n[0], *x))
            t += DT
    if not quiet:
        print(f"elapsed: {time.perf_counter() - now:.2f} sec")
    mppi.close()
    return np.array(rows)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--nonlinear", action="store_true")
    ap.add_argument("--samples", type=int, default=800_000)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--csv", default="logs/mppi/mppi.csv")
    a = ap.parse_args()
    run(a.nonlinear, a.samples, a.seconds, a.csv)
