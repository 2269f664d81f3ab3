"""Profiling target: a few device-resident MPPI steps of the BASELINE config #4 rollout shape (model NL6, 4096 controllers x
8192 samples x H = 8) in the given precision.   python tools/prof_mppi_cl.py [f64fast|f64|f32] [controllers]"""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from dev_time_mppi import dev_alloc
import ctypes as C
from mpc_rs_b200 import Mppi, models
from mpc_rs_b200 import _abi as A

prec = sys.argv[1] if len(sys.argv) > 1 else "f64fast"
Cn = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
m = Mppi(8, 8192, model=models.NL6, lam=1.4, std_dev=4.0, limit=(-10, 10), precision=prec, dt=0.15, controllers=Cn)
x = np.tile(np.array([0.0, 0.0, 0.05, 0.0]), (Cn, 1))
u = np.zeros((Cn, 8))
d_x, d_u, d_o = dev_alloc(x.nbytes), dev_alloc(u.nbytes), dev_alloc(u.nbytes)
A.lib().mpcb_device_upload(0, d_x, x.ctypes.data_as(C.c_void_p), x.nbytes)
A.lib().mpcb_device_upload(0, d_u, u.ctypes.data_as(C.c_void_p), u.nbytes)
for _ in range(4):
    m.compute_device(d_x, d_u, d_o)
m.sync()
print("ok", m.last_info()[0])
