"""Developer probe: launch-to-launch intervals of back-to-back MPPI steps, one CUDA event between consecutive launches."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from mpc_rs_b200 import Mppi, models
from mpc_rs_b200 import _abi as A
import dev_ws
lib = C.CDLL(os.path.join("tools", "libmpcb_e2e.so"))
lib.mpcb_device_loop_events.restype = C.c_int
lib.mpcb_device_loop_events.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float)]
K = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
H = int(sys.argv[2]) if len(sys.argv) > 2 else 100
for ws in sys.argv[3:] or ["-1", "1"]:
    os.environ["MPCB_MPPI_WS"] = ws
    os.environ.setdefault("MPCB_MPPI_WS_CQ", "5")
    m = Mppi(H, K, model=models.NL, lam=0.5, std_dev=3.0, limit=(-20, 20), precision="f32", dt=0.8 / H)
    d_x, d_u, d_o = dev_ws.dev_alloc(32), dev_ws.dev_alloc(8 * H), dev_ws.dev_alloc(8 * H)
    x0, u0 = np.array([0.5, 0, 0.1, 0.0]), np.zeros(H)
    A.lib().mpcb_device_upload(0, d_x, x0.ctypes.data_as(C.c_void_p), 32)
    A.lib().mpcb_device_upload(0, d_u, u0.ctypes.data_as(C.c_void_p), 8 * H)
    n = 1000
    ms = (C.c_float * n)()
    for rep in range(2):
        st = lib.mpcb_device_loop_events(m._h, m.stream, d_x, d_u, d_o, n, ms)
    a = np.array(ms[100:]) * 1e3
    print(f"K={K} H={H} ws={ws}: status {st} interval us: mean {a.mean():.2f} median {np.median(a):.2f} p10 {np.percentile(a,10):.2f} p90 {np.percentile(a,90):.2f} min {a.min():.2f}", flush=True)
    m.close()
