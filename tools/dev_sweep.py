"""Developer probe: device-resident MPPI step time over K / H, to separate per-warp latency, issue limits and tail."""
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


def run(model, H, K, dt, prec="f32", reps=200, C_=1):
    m = Mppi(H, K, model=model, lam=0.5, std_dev=3.0, limit=(-20, 20), precision=prec, dt=dt, controllers=C_)
    x = np.tile(np.array([0.5, 0, 0.1, 0.0]), (C_, 1))
    u = np.zeros((C_, H))
    d_x, d_u, d_o = dev_alloc(x.nbytes), dev_alloc(u.nbytes), dev_alloc(u.nbytes)
    A.lib().mpcb_device_upload(0, d_x, x.ctypes.data_as(C.c_void_p), x.nbytes)
    A.lib().mpcb_device_upload(0, d_u, u.ctypes.data_as(C.c_void_p), u.nbytes)
    for _ in range(10):
        m.compute_device(d_x, d_u, d_o)
    m.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        m.compute_device(d_x, d_u, d_o)
    m.sync()
    t = (time.perf_counter() - t0) / reps
    print(f"{model.name} {prec} C={C_} K={K:8d} H={H:4d}: {t*1e6:8.1f} us  {C_*K*H/t:.3e} steps/s  ({K//128} blocks)", flush=True)
    m.close()


if __name__ == "__main__":
    for K in (128 * 148, 128 * 148 * 2, 128 * 148 * 3, 65536, 128 * 148 * 4, 128 * 148 * 8, 128 * 148 * 16):
        run(models.NL, 100, K, 0.008)
    for H in (4, 8, 20, 52, 100, 200):
        run(models.NL, H, 65536, 0.8 / H)
    for K in (65536, 1 << 18, 1 << 20, 1 << 22):
        run(models.NL, 200, K, 0.004, reps=20)
    run(models.L, 8, 800000, 0.1)
    run(models.NL, 8, 800000, 0.1)
