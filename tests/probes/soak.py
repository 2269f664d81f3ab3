"""Developer soak: random MPPI shapes (K, H, controllers, model, precision, forced kernel flavours) in replay mode against
the oracle.  Not a test (too slow on the oracle side for the suite); run on a GPU box after touching the launch plan."""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import oracle_lib as O
from mpc_rs_b200 import Mppi, models

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
MODELS = [(models.L, O.MODEL_L, 0.5, 3.0, (-20.0, 20.0)), (models.NL, O.MODEL_NL, 0.5, 3.0, (-20.0, 20.0)),
          (models.NL6, O.MODEL_NL6, 1.4, 4.0, (-10.0, 10.0))]
bad = 0
for it in range(n):
    model, oid, lam, sig, lim = MODELS[rng.integers(0, 3)]
    H = int(rng.choice([1, 2, 3, 5, 8, 13, 31, 32, 33, 64, 100, 127, 200, 257, 512]))
    C = int(rng.choice([1, 1, 1, 2, 3, 7, 33, 150]))
    budget = 3_000_000  # oracle rollout-steps per case
    Kmax = max(1, budget // (H * C))
    K = int(min(Kmax, rng.choice([1, 31, 32, 33, 100, 1000, 4096, 20000, 65536, 75777, 200000, 300001])))
    prec = "f64" if rng.random() < 0.4 else "f32"
    env = {}
    if prec == "f32" and rng.random() < 0.5:
        env["MPCB_MPPI_SPT"] = str(rng.integers(1, 3))
    if rng.random() < 0.3:
        env["MPCB_MPPI_VT"] = str(rng.integers(0, 2))
    for k in ("MPCB_MPPI_SPT", "MPCB_MPPI_VT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    dt = float(rng.choice([0.004, 0.01, 0.05])) if oid != O.MODEL_NL6 else 0.02
    p = O.model_defaults(oid, dt=dt)
    xs = rng.normal(0, 0.1, (C, 4))
    us = rng.uniform(-1, 1, (C, H))
    eps = sig * rng.standard_normal((C, K, H))
    tol = 1e-9 if prec == "f64" else (2e-5 if oid != O.MODEL_NL6 else 1e-3)
    try:
        with Mppi(H, K, model=model, lam=lam, std_dev=sig, limit=lim, precision=prec, dt=dt, controllers=C) as m:
            u = m.compute_replay(xs, us, eps.astype(np.float32) if prec == "f32" else eps)
            worst = 0.0
            for c in range(C):
                st, uo, io, _ = O.mppi_compute(oid, p, K, H, lam, sig, lim[0], lim[1], xs[c], us[c],
                                               eps[c].astype(np.float32).astype(np.float64) if prec == "f32" else eps[c])
                assert st == 0 and m.info[c]["status"] == 0
                assert m.info[c]["argmax"] == io["argmax"], (c, m.info[c], io)
                assert m.info[c]["n_finite"] == io["n_finite"]
                worst = max(worst, np.linalg.norm(u[c] - uo) / max(np.linalg.norm(uo), 1e-300))
            ok = worst < tol
    except Exception as e:  # noqa: BLE001
        ok, worst = False, repr(e)[:200]
    bad += 0 if ok else 1
    print(f"{'ok ' if ok else 'BAD'} {model.name:3s} {prec} C={C:3d} K={K:6d} H={H:3d} dt={dt} env={env} worst={worst}", flush=True)
print("soak failures:", bad)
sys.exit(1 if bad else 0)
