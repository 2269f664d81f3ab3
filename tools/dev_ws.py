"""Developer probe for the warp-specialised MPPI kernels (mppi_ws_kernel.cuh): for each variant of kWsVariants
(MPCB_MPPI_WS, -1 = the one-thread-per-sample kernels) replay parity against the oracle and the back-to-back
device-resident step time.   python tools/dev_ws.py [shape] [variants...]      shape: c1 | big | h8"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import oracle_lib as O
from mpc_rs_b200 import Mppi, models
from mpc_rs_b200 import _abi as A

SHAPES = {"c1": (100, 65536, 0.008, 400), "big": (200, 1 << 20, 0.004, 10), "h8": (8, 800000, 0.1, 50), "mid": (100, 1 << 18, 0.008, 20)}


def dev_alloc(nbytes):
    p = C.c_void_p()
    A.check(A.lib().mpcb_device_alloc(0, nbytes, C.byref(p)))
    return p.value


_E2E = None


def device_loop(m, d_x, d_u0, d_u1, n):
    """n back-to-back steps enqueued from compiled code (tools/e2e_loop.c): a Python loop is host-bound at ~30 us/step."""
    global _E2E
    if _E2E is None:
        _E2E = C.CDLL(os.path.join("tools", "libmpcb_e2e.so"))
        _E2E.mpcb_device_loop.restype = C.c_int
        _E2E.mpcb_device_loop.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    st = _E2E.mpcb_device_loop(m._h, d_x, d_u0, d_u1, n)
    if st != 0:
        raise RuntimeError(f"device loop status {st}")


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "c1"
    if shape.startswith("K"):  # K<samples>xH<horizon>
        ks, hs = shape[1:].split("xH")
        SHAPES[shape] = (int(hs), int(ks), 0.8 / int(hs), 200)
    H, K, dt, reps = SHAPES[shape]
    variants = [int(v) for v in sys.argv[2:]] or [-1, 0, 1, 2, 3]
    cqs = [int(v) for v in os.environ.get("CQS", "2").split(",")]
    lam, sig, lim = 0.5, 3.0, (-20.0, 20.0)
    x0, u0 = np.array([0.5, 0.0, 0.1, 0.0]), np.zeros(H)
    Kp = min(K, 65536)
    rng = np.random.default_rng(1)
    eps = (sig * rng.standard_normal((Kp, H))).astype(np.float32)
    p = O.model_defaults(O.MODEL_NL, dt=dt)
    st, u_ref, info_ref, _ = O.mppi_compute(O.MODEL_NL, p, Kp, H, lam, sig, lim[0], lim[1], x0, u0, eps.astype(np.float64))
    u1 = u_ref.copy()
    st, u_ref2, info_ref2, _ = O.mppi_compute(O.MODEL_NL, p, Kp, H, lam, sig, lim[0], lim[1], x0, u1, eps.astype(np.float64))
    for v in variants:
        for cq in (cqs if v >= 0 else [0]):
            os.environ["MPCB_MPPI_WS"] = str(v)
            os.environ["MPCB_MPPI_WS_CQ"] = str(max(cq, 1))
            os.environ["MPCB_MPPI_WS_DEBUG"] = os.environ.get("WSDBG", "0")
            try:
                with Mppi(H, Kp, model=models.NL, lam=lam, std_dev=sig, limit=lim, precision="f32", dt=dt) as m:
                    u = m.compute_replay(x0, u0, eps)
                    e1 = np.linalg.norm(u - u_ref) / np.linalg.norm(u_ref)
                    a1 = m.info[0]["argmax"] == info_ref["argmax"]
                    u = m.compute_replay(x0, u1, eps)
                    e2 = np.linalg.norm(u - u_ref2) / np.linalg.norm(u_ref2)
                    a2 = m.info[0]["argmax"] == info_ref2["argmax"]
                with Mppi(H, K, model=models.NL, lam=lam, std_dev=sig, limit=lim, precision="f32", dt=dt) as m:
                    d_x, d_u, d_o = dev_alloc(32), dev_alloc(8 * H), dev_alloc(8 * H)
                    A.lib().mpcb_device_upload(0, d_x, x0.ctypes.data_as(C.c_void_p), 32)
                    A.lib().mpcb_device_upload(0, d_u, u0.ctypes.data_as(C.c_void_p), 8 * H)
                    device_loop(m, d_x, d_u, d_o, 6)
                    m.sync()
                    best = 1e9
                    for _ in range(3):
                        t0 = time.perf_counter()
                        device_loop(m, d_x, d_u, d_o, reps)
                        m.sync()
                        best = min(best, (time.perf_counter() - t0) / reps)
                    ug = m.compute(x0, u0)
                    print(f"{shape} variant {v:2d} cq {cq}: replay err {e1:.2e} {e2:.2e} argmax {a1} {a2} | device {best*1e6:8.2f} us "
                          f"({K*H/best:.3e} steps/s) frac {60*K*H/best/69.1e12:.3f} | generate finite {bool(np.all(np.isfinite(ug)))}", flush=True)
            except Exception as ex:  # noqa: BLE001
                print(f"{shape} variant {v} cq {cq}: FAILED {type(ex).__name__}: {ex}", flush=True)


if __name__ == "__main__":
    main()
