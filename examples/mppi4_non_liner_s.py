#!/usr/bin/env python
"""examples/mppi4-non-liner-s.rs on B200: MPPI (model NL, K = 1 500 000, N = 8, R = 10, limit +-10) steering the nonlinear
pendulum from the estimate of the library UKF (mpc::ukf, n = 4, o = 3, run-time dt), with the reference's three
free-running threads (plant 1 ms :51-63, sensor + UKF ~3 ms :65-118, MPPI as fast as it computes :120-173) replaced by
a fixed schedule on a simulated clock: plant every 1 ms, sensor -> predict(dt) -> update every 3 ms, MPPI every
`--period` seconds (10 ms).  Prints the reference's "Rcv:" / "Con:" lines and writes logs/mppi/mppi.csv in its format
(t, u, x_est[0..4], :157-166).

    python examples/mppi4_non_liner_s.py [--samples K] [--seconds 5] [--period 0.01] [--csv logs/mppi/mppi.csv]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_rs_b200 import Mppi, UnscentedKalmanFilter, models  # noqa: E402
from mpc_rs_b200.csvlog import MppiLog  # noqa: E402
from mpc_rs_b200.plants import PlantNL, PlantPenNL  # noqa: E402

T, N = 0.8, 8                                   # :12-14
DT = T / N
K, LAMBDA, R, LIMIT = 1_500_000, 0.5, 10.0, (-10.0, 10.0)  # :17-22
P0 = np.eye(4)                                   # :218-223
Q = np.array([[0, 0, 0, 0], [0, 0, 0, 1.0], [0, 0, 1.0, 1e2], [0, 1.0, 1e2, 1e4]])  # :224-229
R_OBS = np.diag([50.0, 50.0, 0.5])               # :230-234
SENSOR_NOISE = np.array([50.0, 50.0, 0.5])       # :241-247


def run(samples=K, seconds=5.0, period=0.01, csv="logs/mppi/mppi.csv", quiet=False, seed=0):
    rng = np.random.default_rng(seed)
    plant_hx = PlantPenNL()                      # hx :250-256 is the o = 3 measurement of ukf-pen2.rs
    x = np.array([0.0, 0.0, 0.01, 0.0])          # :34
    u_n = np.zeros(N)
    mppi = Mppi.new(models.NL, models.NL, LAMBDA, R, LIMIT, N=N, K=samples, seed=seed, dt=DT)
    ukf = UnscentedKalmanFilter.new(x, P0, Q, R_OBS, fx=models.PEN_NL)
    rows = []
    plant_dt, ukf_every, mppi_every = 1e-3, 3, max(1, int(round(period / 1e-3)))
    with MppiLog(csv) as log:
        for k in range(int(round(seconds / plant_dt))):
            t = k * plant_dt
            x = PlantNL(plant_dt).step(x, u_n[0])                     # dynamics_short(&x, u_n[0], dt) :57
            if k % ukf_every == 0:
                z = plant_hx.hx(x) + SENSOR_NOISE * rng.standard_normal(3)
                ukf.predict(u_n[0], models.PEN_NL, dt=ukf_every * plant_dt)  # fx = |x, u| dynamics_short(x, u, dt) :88-89
                ukf.update(z, models.PEN_NL)
                if not quiet:
                    e, p = ukf.state(), np.diag(ukf.covariance())
                    print(f"Rcv: t: {t:.2f} x: [{x[0]:6.2f}, {x[1]:5.2f}, {x[2]:5.2f}, {x[3]:5.2f}] "
                          f"est: [{e[0]:6.2f}, {e[1]:5.2f}, {e[2]:5.2f}, {e[3]:5.2f}] "
                          f"p: [{p[0]:6.2f}, {p[1]:5.2f}, {p[2]:5.2f}, {p[3]:5.2f}] ")
            if k % mppi_every == 0:
                x_est = ukf.state()
                if abs(x_est[2]) > np.radians(60.0):                  # :123-126
                    print("x[2] is over 60 degrees")
                    break
                try:
                    u_n = mppi.compute(x_est, u_n)
                    err = None
                except Exception as e:                                # Err(e) => zeros :133-136
                    u_n, err = np.zeros(N), e
                if not quiet:
                    print(f"Con: t: {t:.2f} est: [{x_est[0]:6.2f}, {x_est[1]:5.2f}, {x_est[2]:5.2f}, {x_est[3]:5.2f}] "
                          f"u: {u_n[0]:8.3f} " + (f"Failed to compute MPPI: {err}" if err else ""))
                log.write(t, u_n[0], x_est)
                rows.append((t, u_n[0], *x_est, *x))
    mppi.close()
    ukf.close()
    return np.array(rows)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=K)
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--period", type=float, default=0.01)
    ap.add_argument("--csv", default="logs/mppi/mppi.csv")
    a = ap.parse_args()
    run(a.samples, a.seconds, a.period, a.csv)
