"""Profiling target: a few device-resident fused UKF steps of BASELINE config #3 (examples/ukf-pen.rs, B = 2^20)."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from mpc_rs_b200 import BatchedUkf, models, ukf
from mpc_rs_b200 import _abi as A

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
name = sys.argv[3] if len(sys.argv) > 3 else "PEN_LIN"
model = getattr(models, name)
f = BatchedUkf(model, B)
Q, R, P0 = ukf.default_noise(model, 0.01)
f.init(np.zeros(f.n), P0, Q, R)
rng = np.random.default_rng(0)
z = np.ascontiguousarray(rng.normal(0, 0.7, (T, f.o, B)))
d_z = C.c_void_p()
A.check(A.lib().mpcb_device_alloc(0, z.nbytes, C.byref(d_z)))
A.check(A.lib().mpcb_device_upload(0, d_z, z.ctypes.data_as(C.c_void_p), z.nbytes))
for t in range(T):
    f.run_device(1, d_z.value + t * f.o * B * 8, u=0.0015, dt=0.01)
f.sync()
f.run_device(T, d_z.value, u=0.0015, dt=0.01)
f.sync()
print("ok failed:", int((f.status() != 0).sum()))
