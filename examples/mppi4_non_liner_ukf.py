#!/usr/bin/env python
"""examples/mppi4-non-liner-ukf.rs on B200, batched (BASELINE config #4): C independent robots, each balanced by its
own MPPI controller fed by its own UKF, on a fixed 10 ms tick (see mpc_rs_b200/closed_loop.py for the schedule).
Robot 0 is logged in the reference's 20-column CSV (t, u, x, x_est, x_pred) every 30 ms like the logging thread.

    python examples/mppi4_non_liner_ukf.py [--controllers 4096] [--samples 8192] [--seconds 3] [--truth]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_rs_b200.closed_loop import ClosedLoopBatch  # noqa: E402
from mpc_rs_b200.csvlog import MppiUkfLog  # noqa: E402


def run(controllers=4096, samples=8192, seconds=3.0, truth=False, csv="logs/mppi/mppi.csv", quiet=False, seed=20240004,
        precision=None):
    rng = np.random.default_rng(seed)
    x0 = np.zeros((controllers, 6))
    x0[:, 3] = rng.uniform(-0.1, 0.1, controllers)  # theta0 ~ U(-0.1, 0.1); the example starts at 0
    x0[0, 3] = 0.0
    with ClosedLoopBatch(controllers, samples, use_estimate=not truth, seed=seed, x0=x0, precision=precision) as loop, \
            MppiUkfLog(csv) as log:
        next_log = 0.0
        loop.sync()
        t0 = time.perf_counter()  # the loop itself: construction (CUDA context, buffers) is not part of a tick
        while loop.t < seconds:
            loop.tick()
            if loop.t >= next_log:  # the logging thread writes every 30 ms (:404)
                next_log += 0.03
                x_est, _ = loop.estimate_range(0, 1)
                u_n = loop.controls()[0]
                x_pred = x_est[0].copy()
                for i in range(loop.H):  # :418-421
                    x_pred = loop.plant.dynamics_short(x_pred, u_n[i], loop.DT, 0.0)
                x_true = loop.x
                log.write(loop.t, u_n[0], x_true[0], x_est[0], x_pred)
                if not quiet:
                    e, x = x_est[0], x_true[0]
                    print(f"t:{loop.t:6.2f} u:{u_n[0]:6.2f} e:[{e[0]:6.2f},{e[1]:6.2f},{np.degrees(e[3]):5.0f},{np.degrees(e[4]):5.0f}] "
                          f"x:[{x[0]:6.2f},{x[1]:6.2f},{np.degrees(x[3]):5.0f},{np.degrees(x[4]):5.0f}] upright {int(loop.upright().sum())}/{controllers}")
        dt = time.perf_counter() - t0
        up = loop.upright()
        ticks = loop.ticks
    if not quiet:
        print(f"{ticks} ticks x {controllers} robots in {dt:.2f} s: {ticks * controllers * samples * 8 / dt:.3e} rollout-steps/s, "
              f"{ticks * controllers / dt:.3e} filter-updates/s end to end; {int(up.sum())} robots upright")
    return up


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--controllers", type=int, default=4096)
    ap.add_argument("--samples", type=int, default=8192)
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--truth", action="store_true", help="DEBUG_UKF = true: feed the controller the true state")
    ap.add_argument("--csv", default="logs/mppi/mppi.csv")
    ap.add_argument("--precision", choices=["f32", "f64", "f64fast"], default=None,
                    help="MPPI arithmetic; default f64fast for this model (DESIGN.md 4.1 precision policy): FP64 folded form, 1e-9 from the reference; f64 = reference order (2x slower), f32 (2x faster) misses 1e-5 on this model")
    a = ap.parse_args()
    run(a.controllers, a.samples, a.seconds, a.truth, a.csv, precision=a.precision)
