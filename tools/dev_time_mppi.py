"""Developer timing probe (not the bench): device-resident and host end-to-end time of one MPPI step."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mpc_rs_b200 import Mppi, models
from mpc_rs_b200 import _abi as A


def dev_alloc(nbytes):
    p = C.c_void_p()
    A.check(A.lib().mpcb_device_alloc(0, nbytes, C.byref(p)))
    return p.value


def run(model, H, K, dt, prec, C_=1, reps=50):
    m = Mppi(H, K, model=model, lam=0.5, std_dev=3.0, limit=(-20, 20), precision=prec, dt=dt, controllers=C_)
    x = np.tile(np.array([0.5, 0, 0.1, 0.0]), (C_, 1))
    u = np.zeros((C_, H))
    d_x, d_u, d_o = dev_alloc(x.nbytes), dev_alloc(u.nbytes), dev_alloc(u.nbytes)
    A.lib().mpcb_device_upload(0, d_x, x.ctypes.data_as(C.c_void_p), x.nbytes)
    A.lib().mpcb_device_upload(0, d_u, u.ctypes.data_as(C.c_void_p), u.nbytes)
    for _ in range(5):
        m.compute_device(d_x, d_u, d_o)
    m.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        m.compute_device(d_x, d_u, d_o)
    m.sync()
    t_dev = (time.perf_counter() - t0) / reps
    for _ in range(5):
        m.compute(x, u)
    t0 = time.perf_counter()
    for _ in range(reps):
        m.compute(x, u)
    t_e2e = (time.perf_counter() - t0) / reps
    steps = C_ * K * H
    print(f"{model.name} {prec} C={C_} K={K} H={H}: device {t_dev*1e6:8.1f} us ({steps/t_dev:.3e} steps/s)   "
          f"host e2e {t_e2e*1e6:8.1f} us ({steps/t_e2e:.3e} steps/s)  block={m.cfg.horizon}", flush=True)
    m.close()


if __name__ == "__main__":
    run(models.NL, 100, 65536, 0.008, "f32")
    run(models.NL, 100, 65536, 0.008, "f64")
    run(models.NL, 200, 1 << 20, 0.004, "f32", reps=10)
    run(models.NL, 200, 1 << 20, 0.004, "f64", reps=5)
    run(models.L, 8, 800000, 0.1, "f32")
    run(models.NL, 8, 800000, 0.1, "f32")
    run(models.NL6, 8, 8192, 0.15, "f32", C_=4096, reps=5)
