"""Developer probe: the six-state fast kernels (streaming, or general with MPCB_UKF_NO_STREAM=1) and the reference-order kernel
against the oracle on random states / covariances (the draws of tests/test_ukf_gpu.py::test_six_state_streaming_kernel_...)."""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import oracle_lib as O
from mpc_rs_b200 import BatchedUkf, models

MODELS = {"PEN6": (models.PEN6, O.MODEL_PEN6, 0.1), "NL6_UKF": (models.NL6_UKF, O.MODEL_NL6_UKF, 0.3),
          "PEN_LIN": (models.PEN_LIN, O.MODEL_PEN_LIN, 0.0015), "PEN_NL": (models.PEN_NL, O.MODEL_PEN_NL, 0.1)}
SQRT = {"cholesky": O.SQRT_CHOLESKY, "eig": O.SQRT_EIG}


def relerr(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / max(np.linalg.norm(np.ravel(b)), 1e-300)


rng = np.random.default_rng(20240611)
for name in (sys.argv[1:] or ["PEN6", "NL6_UKF"]):
    model, oid, u = MODELS[name]
    p = O.model_defaults(oid)
    n, o = O.dims(oid)
    dt = 0.01 if name == "NL6_UKF" else 0.0
    for sqrt_mode in ("eig", "cholesky"):
        for trial in range(4):
            B = int(rng.integers(1, 300))
            a = rng.normal(0, 1, (B, n, n))
            scale = float(np.exp(rng.uniform(np.log(1e-3), np.log(50.0))))
            P = scale * (a @ np.transpose(a, (0, 2, 1)) + 0.05 * np.eye(n))
            if sqrt_mode == "cholesky" and trial % 2 == 1:
                rng.integers(0, B)
            x = rng.normal(0, 0.2, (B, n))
            Q = np.diag(rng.uniform(0, 1, n))
            R = np.diag(rng.uniform(0.05, 5.0, o))
            z = rng.normal(0, 1, (B, o))
            xr, Pr, st_o = O.ukf_step_batch(oid, p, x, P, Q, R, u, z, dt, SQRT[sqrt_mode], O.ORDER_LIBRARY)
            out = []
            for exact in (False, True):
                with BatchedUkf(model, B, sqrt_mode=sqrt_mode, sigma_order="library", exact=exact) as f:
                    f.init(np.zeros(n), np.eye(n), Q, R)
                    f.set_state(x, P)
                    f.step(u, z, dt, check=False)
                    xg, Pg = f.get_state()
                out.append((relerr(xg, xr), relerr(Pg, Pr)))
            # per-filter worst case of the fast kernel
            print(f"{name:8s} {sqrt_mode:8s} B={B:3d} scale={scale:8.3g}: fast x {out[0][0]:.1e} P {out[0][1]:.1e} | exact x {out[1][0]:.1e} P {out[1][1]:.1e}", flush=True)
