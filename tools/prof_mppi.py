"""Profiling target: a few device-resident MPPI steps of BASELINE config #2 (model NL, K=65536, H=100, FP32)."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from mpc_rs_b200 import Mppi, models
from mpc_rs_b200 import _abi as A

K = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
H = int(sys.argv[2]) if len(sys.argv) > 2 else 100
prec = sys.argv[3] if len(sys.argv) > 3 else "f32"
n = int(sys.argv[4]) if len(sys.argv) > 4 else 6


def dev_alloc(nbytes):
    p = C.c_void_p()
    A.check(A.lib().mpcb_device_alloc(0, nbytes, C.byref(p)))
    return p.value


m = Mppi(H, K, model=models.NL, lam=0.5, std_dev=3.0, limit=(-20, 20), precision=prec, dt=0.8 / H)
x = np.array([[0.5, 0, 0.1, 0.0]])
u = np.zeros((1, H))
d_x, d_u, d_o = dev_alloc(x.nbytes), dev_alloc(u.nbytes), dev_alloc(u.nbytes)
A.lib().mpcb_device_upload(0, d_x, x.ctypes.data_as(C.c_void_p), x.nbytes)
A.lib().mpcb_device_upload(0, d_u, u.ctypes.data_as(C.c_void_p), u.nbytes)
for _ in range(n):
    m.compute_device(d_x, d_u, d_o)
m.sync()
print("ok", m.last_info()[0])
