"""Developer probe: what a user-supplied model costs against the built-in one (BASELINE configs[1] shape, model NL).
The user source is the straightforward port of examples/mppi4-non-liner.rs (tests/test_user_model_gpu.py); a second
variant uses the library helpers (mpcb::sincos_r, mpcb::fast_rcp) the built-in FP32 model is written with."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from mpc_rs_b200 import Mppi, models, user_model  # noqa: E402
from mpc_rs_b200 import _abi as A  # noqa: E402

PORT = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "test_user_model_gpu.py")).read()
PORT = PORT.split('NL_SOURCE = r"""')[1].split('"""')[0]
TUNED = PORT.replace("const real s = sin(x[2]), c = cos(x[2]);", "real s, c; mpcb::sincos_r(x[2], &s, &c);")


def dev_alloc(n):
    p = C.c_void_p()
    A.check(A.lib().mpcb_device_alloc(0, n, C.byref(p)))
    return p.value


def timeit(m, H, reps=200):
    x = np.array([[0.5, 0, 0.1, 0.0]])
    u = np.zeros((1, H))
    d_x, d_u, d_o = dev_alloc(32), dev_alloc(8 * H), dev_alloc(8 * H)
    A.lib().mpcb_device_upload(0, d_x, x.ctypes.data_as(C.c_void_p), 32)
    A.lib().mpcb_device_upload(0, d_u, u.ctypes.data_as(C.c_void_p), 8 * H)
    best = 1e9
    for _ in range(3):
        for _ in range(5):
            m.compute_device(d_x, d_u, d_o)
        m.sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            m.compute_device(d_x, d_u, d_o)
        m.sync()
        best = min(best, (time.perf_counter() - t0) / reps)
    return best


if __name__ == "__main__":
    for K, H, dt in ((65536, 100, 0.008), (1 << 20, 200, 0.004)):
        prm = [150e-3, 50e-3, 2.3 - 2.0 * 150e-3 + 2.0, 0.2474, 150e-3 * 50e-3 * 50e-3, 0.2, 9.81, 0.15, dt]
        kw = dict(lam=0.5, std_dev=3.0, limit=(-20, 20), precision="f32")
        os.environ["MPCB_MPPI_SPT"] = "1"
        rows = [("built-in NL, scalar kernels", Mppi(H, K, model=models.NL, dt=dt, **kw))]
        del os.environ["MPCB_MPPI_SPT"]
        rows.append(("built-in NL, plan as shipped", Mppi(H, K, model=models.NL, dt=dt, **kw)))
        t0 = time.perf_counter()
        rows.append(("user port (sin, cos, /)", Mppi(H, K, model=user_model(PORT, prm), **kw)))
        t_compile = time.perf_counter() - t0
        rows.append(("user port + mpcb::sincos_r", Mppi(H, K, model=user_model(TUNED, prm), **kw)))
        print(f"K={K} H={H}  (create_user took {t_compile:.1f} s)")
        for name, m in rows:
            t = timeit(m, H, reps=200 if K < 1e6 else 10)
            print(f"  {name:32s} {t * 1e6:9.1f} us  {K * H / t:.3e} rollout-steps/s")
            m.close()
