import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from mpc_rs_b200 import Mppi, models
from mpc_rs_b200 import _abi as A
import dev_ws
for K, H in ((1024, 8), (65536, 100)):
    for ws in ("-1", "1"):
        os.environ["MPCB_MPPI_WS"] = ws
        m = Mppi(H, K, model=models.NL, lam=0.5, std_dev=3.0, limit=(-20, 20), precision="f32", dt=0.8 / H)
        d_x, d_u, d_o = dev_ws.dev_alloc(32), dev_ws.dev_alloc(8 * H), dev_ws.dev_alloc(8 * H)
        x0, u0 = np.array([0.5, 0, 0.1, 0.0]), np.zeros(H)
        A.lib().mpcb_device_upload(0, d_x, x0.ctypes.data_as(C.c_void_p), 32)
        A.lib().mpcb_device_upload(0, d_u, u0.ctypes.data_as(C.c_void_p), 8 * H)
        dev_ws.device_loop(m, d_x, d_u, d_o, 10); m.sync()
        n = 2000
        t0 = time.perf_counter(); dev_ws.device_loop(m, d_x, d_u, d_o, n); t1 = time.perf_counter(); m.sync(); t2 = time.perf_counter()
        print(f"K={K} H={H} ws={ws}: enqueue {1e6*(t1-t0)/n:.2f} us/launch, total {1e6*(t2-t0)/n:.2f} us/launch", flush=True)
        m.close()
