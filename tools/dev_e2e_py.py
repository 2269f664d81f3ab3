import sys, time
import numpy as np
sys.path.insert(0, ".")
from mpc_rs_b200 import Mppi, models
m = Mppi(100, 65536, model=models.NL, lam=0.5, std_dev=3.0, limit=(-20, 20), precision="f32", dt=0.008)
x = np.array([0.5, 0, 0.1, 0.0]); u = np.zeros(100)
for _ in range(50): u = m.compute(x, u)
t0 = time.perf_counter()
for _ in range(2000): u = m.compute(x, u)
print(f"python Mppi.compute: {(time.perf_counter()-t0)/2000*1e6:.2f} us/call")
xb, ub, out, infos, px, pu, po, fn = m._fast_path()
t0 = time.perf_counter()
for _ in range(2000): fn(m._h, px, pu, po, infos)
print(f"python raw ctypes call: {(time.perf_counter()-t0)/2000*1e6:.2f} us/call")
