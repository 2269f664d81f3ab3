"""Developer timing probe: device-resident fused UKF steps for every model (one launch per step and T steps fused)."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mpc_rs_b200 import BatchedUkf, models, ukf
from mpc_rs_b200 import _abi as A

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
T = 20
for name, exact in (("PEN_LIN", False), ("PEN_LIN", True), ("PEN_NL", False), ("PEN6", False), ("NL6_UKF", False), ("NL6_UKF", True)):
    model = getattr(models, name)
    f = BatchedUkf(model, B, exact=exact)
    Q, R, P0 = ukf.default_noise(model, 0.01)
    f.init(np.zeros(f.n), P0, Q, R)
    rng = np.random.default_rng(0)
    z = np.ascontiguousarray(rng.normal(0, 0.7, (T, f.o, B)))
    d_z = C.c_void_p()
    A.check(A.lib().mpcb_device_alloc(0, z.nbytes, C.byref(d_z)))
    A.check(A.lib().mpcb_device_upload(0, d_z, z.ctypes.data_as(C.c_void_p), z.nbytes))
    for t in range(3):
        f.run_device(1, d_z.value + t * f.o * B * 8, u=0.0015, dt=0.01)
    f.sync()
    t0 = time.perf_counter()
    for t in range(T):
        f.run_device(1, d_z.value + t * f.o * B * 8, u=0.0015, dt=0.01)
    f.sync()
    dt1 = (time.perf_counter() - t0) / T
    f.init(np.zeros(f.n), P0, Q, R)
    t0 = time.perf_counter()
    f.run_device(T, d_z.value, u=0.0015, dt=0.01)
    f.sync()
    dtT = (time.perf_counter() - t0) / T
    bytes_per = 8 * (2 * f.n + 2 * f.n * f.n + f.o)
    print(f"{name:8s} exact={int(exact)} n={f.n} o={f.o}: per-step launch {dt1*1e6:8.1f} us = {B/dt1:.3e} upd/s = {B*bytes_per/dt1/1e9:7.0f} GB/s algorithmic; "
          f"fused {dtT*1e6:8.1f} us/step = {B/dtT:.3e} upd/s; failed {int((f.status()!=0).sum())}", flush=True)
    A.lib().mpcb_device_free(0, d_z)
    f.close()
