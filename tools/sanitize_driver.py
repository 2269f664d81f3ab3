"""tools/sanitize_driver.py — small MPPI/UKF calls through the C ABI for compute-sanitizer (developer probe).

    compute-sanitizer --tool memcheck  python tools/sanitize_driver.py
    compute-sanitizer --tool racecheck python tools/sanitize_driver.py
    compute-sanitizer --tool synccheck python tools/sanitize_driver.py

Every kernel flavour is hit once at a small shape: scalar / packed / no-v-tile FP32, FP64, replay, batched
controllers, merge tree, the peer exchange between two handles of this process, and the UKF kernels.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_rs_b200 import models  # noqa: E402
from mpc_rs_b200.mppi import Mppi  # noqa: E402
from mpc_rs_b200.ukf import BatchedUkf, default_noise  # noqa: E402

X0 = np.array([0.5, 0.0, 0.1, 0.0])


def mppi_case(tag, H, K, model, prec, controllers=1, env=None, replay=False, steps=2):
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        with Mppi(H, K, model=model, lam=0.5, std_dev=3.0, limit=(-20, 20), precision=prec, controllers=controllers,
                  dt=0.05, seed=3) as m:
            u = np.zeros((controllers, H))
            x = np.tile(X0, (controllers, 1))
            for _ in range(steps):
                if replay:
                    eps = (3.0 * np.random.default_rng(1).standard_normal((controllers, K, H))).astype(np.float32)
                    u = np.reshape(m.compute_replay(x, u, eps), (controllers, H))
                else:
                    u = np.reshape(m.compute(x if controllers > 1 else X0, u if controllers > 1 else u[0]), (controllers, H))
            assert np.all(np.isfinite(u)), tag
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    print("ok", tag, flush=True)


def peer_case():
    G, H, K = 2, 12, 6000
    hs = [Mppi(H, K, model=models.NL, lam=0.5, std_dev=3.0, limit=(-20, 20), precision="f32", dt=0.05, rank=r,
               world_size=G, seed=5) for r in range(G)]
    handles = [h.peer_handle() for h in hs]
    for h in hs:
        h.attach_peers(handles)
    import ctypes as C
    from mpc_rs_b200 import _abi as A
    d = [C.c_void_p() for _ in range(2 + G)]
    for q, n in zip(d, [32, 8 * H] + [8 * H] * G):
        A.check(A.lib().mpcb_device_alloc(0, n, C.byref(q)))
    u = np.zeros(H)
    A.check(A.lib().mpcb_device_upload(0, d[0], X0.ctypes.data_as(C.c_void_p), 32))
    A.check(A.lib().mpcb_device_upload(0, d[1], u.ctypes.data_as(C.c_void_p), 8 * H))
    for _ in range(3):
        for r, h in enumerate(hs):
            h.compute_device(d[0].value, d[1].value, d[2 + r].value)
        for h in hs:
            h.sync()
            assert h.last_info()[0]["status"] == 0
    for h in hs:
        h.close()
    for q in d:
        A.lib().mpcb_device_free(0, q)
    print("ok peer exchange", flush=True)


def ukf_case(model, sqrt_mode, B=2048, steps=3):
    with BatchedUkf(model, B, sqrt_mode=sqrt_mode, dt=0.01) as f:
        Q, R, P0 = default_noise(model, 0.01)
        f.init(np.zeros(f.n), P0, Q, R)
        rng = np.random.default_rng(2)
        for i in range(steps):
            z = rng.standard_normal((B, f.o)) * 0.1
            if i == 0:
                f.predict(0.0015, dt=0.01)
                f.update(z)
            else:
                f.step(0.0015, z, dt=0.01)
        x, P = f.get_state()
        assert np.all(np.isfinite(x)) and np.all(np.isfinite(P)), (model.name, sqrt_mode)
    print("ok ukf", model.name, sqrt_mode, flush=True)


if __name__ == "__main__":
    nospin = {"MPCB_MPPI_NO_SPIN": "1"}
    mppi_case("f32 scalar single-batch (spin)", 16, 148 * 64, models.NL, "f32")
    mppi_case("f32 scalar, no spin", 16, 20000, models.NL, "f32", env=nospin)
    mppi_case("f32 packed", 16, 148 * 512, models.NL, "f32", env={"MPCB_MPPI_SPT": "2", **nospin})
    mppi_case("f32 L odd horizon", 7, 5000, models.L, "f32", env=nospin)
    mppi_case("f64 NL6", 8, 5000, models.NL6, "f64", env=nospin)
    mppi_case("f32 replay", 16, 4096, models.NL, "f32", replay=True, env=nospin)
    short = {"MPCB_MPPI_SHORT": "1", **nospin}
    mppi_case("f64fast short-horizon kernel", 8, 20000, models.NL6, "f64fast", env=short)
    mppi_case("f64 short-horizon kernel, odd horizon, batched", 5, 3001, models.NL, "f64", controllers=7, env=short)
    mppi_case("f64fast short-horizon kernel, replay", 8, 4096, models.L, "f64fast", replay=True, env=short)
    mppi_case("f32 batched controllers", 8, 2048, models.NL6, "f32", controllers=24, env=nospin)
    mppi_case("f32 batched packed", 8, 8192, models.NL6, "f32", controllers=8, env={"MPCB_MPPI_SPT": "2", **nospin})
    mppi_case("f32 no v tile", 40, 4096, models.NL, "f32", controllers=8, env={"MPCB_MPPI_VT": "0", **nospin})
    mppi_case("f32 merge tree", 8, 148 * 128 * 20, models.L, "f32", env={"MPCB_MPPI_BLOCK": "128", **nospin}, steps=1)
    peer_case()
    if os.environ.get("SAN_SKIP_UKF") != "1":
        ukf_case(models.PEN_LIN, "cholesky")
        ukf_case(models.PEN_NL, "svd", B=512, steps=2)
        ukf_case(models.NL6_UKF, "svd", B=256, steps=2)
    print("sanitize driver finished", flush=True)
