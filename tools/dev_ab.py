"""Developer A/B probe: device-resident MPPI step time of several library builds (tools/build_variant.sh) on the same box.
    python tools/dev_ab.py base rrr ...      # 'base' = the regular libmpc_b200.so, others = libmpc_b200_<name>.so
"""
import os
import subprocess
import sys

CHILD = r'''
import ctypes as C, sys, time
import numpy as np
sys.path.insert(0, ".")
from mpc_rs_b200 import Mppi, models
from mpc_rs_b200 import _abi as A
def dev_alloc(n):
    p = C.c_void_p(); A.check(A.lib().mpcb_device_alloc(0, n, C.byref(p))); return p.value
out = []
for (H, K, dt, reps) in ((100, 65536, 0.008, 300), (200, 1 << 21, 0.004, 10), (8, 800000, 0.1, 100)):
    m = Mppi(H, K, model=models.NL, lam=0.5, std_dev=3.0, limit=(-20, 20), precision="f32", dt=dt)
    x = np.array([[0.5, 0, 0.1, 0.0]]); u = np.zeros((1, H))
    d_x, d_u, d_o = dev_alloc(32), dev_alloc(8 * H), dev_alloc(8 * H)
    A.lib().mpcb_device_upload(0, d_x, x.ctypes.data_as(C.c_void_p), 32)
    A.lib().mpcb_device_upload(0, d_u, u.ctypes.data_as(C.c_void_p), 8 * H)
    best = 1e9
    for rep in range(4):
        for _ in range(5): m.compute_device(d_x, d_u, d_o)
        m.sync(); t0 = time.perf_counter()
        for _ in range(reps): m.compute_device(d_x, d_u, d_o)
        m.sync(); best = min(best, (time.perf_counter() - t0) / reps)
    out.append("K=%d H=%d: %.2f us (%.3e steps/s)" % (K, H, best * 1e6, K * H / best))
    m.close()
print(" | ".join(out))
'''

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for rnd in range(2):
    for name in sys.argv[1:]:
        env = dict(os.environ)
        if name != "base":
            env["MPCB_LIB_PATH"] = os.path.join(root, "mpc_rs_b200", "libmpc_b200_%s.so" % name)
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, cwd=root, capture_output=True, text=True, timeout=600)
        print("%-8s %s %s" % (name, r.stdout.strip(), r.stderr.strip()[-300:]), flush=True)
