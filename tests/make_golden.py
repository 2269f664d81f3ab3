"""Generates tests/golden/*.npz from the CPU oracle (seeded PCG64 inputs).  The reference ships no golden
vectors, so these fixtures pin the ORACLE: `python tests/make_golden.py` rewrites them, the CPU
suite compares a fresh oracle run against the committed files, and the GPU suite compares the CUDA path
against them without needing the oracle's intermediate state."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_lib as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
X0 = np.array([0.5, 0.0, 0.1, 0.0])


def _mppi(oid, H, dt, lam, sig, lim, K=2048, steps=3, seed=20240001):
    def run():
        p = O.model_defaults(oid, dt=dt)
        rng = np.random.Generator(np.random.PCG64(seed))
        x, u = X0.copy(), np.zeros(H)
        out = dict(H=np.array(H), K=np.array(K), dt=np.array(dt), lam=np.array(lam), sig=np.array(sig), lim=np.array(lim))
        for s in range(steps):
            eps = (sig * rng.standard_normal((K, H))).astype(np.float32)  # f32 noise: usable by both precisions
            out[f"x_{s}"], out[f"u_in_{s}"], out[f"eps_{s}"] = x.copy(), u.copy(), eps
            st, u, info, c = O.mppi_compute(oid, p, K, H, lam, sig, lim[0], lim[1], x, u, eps.astype(np.float64), want_costs=True)
            assert st == 0
            out[f"u_out_{s}"], out[f"argmax_{s}"], out[f"max_{s}"], out[f"sum_{s}"] = u.copy(), np.array(info["argmax"]), np.array(info["max"]), np.array(info["sum"])
            x = O.dynamics(oid, p, x, u[0])
        return out
    return run


def _ukf(oid, sqrt_mode, order, u, dt=0.0, B=16, T=5, seed=20240003):
    def run():
        p = O.model_defaults(oid)
        n, o = O.dims(oid)
        Q, R, P0 = O.ukf_default_noise(oid, dt)
        rng = np.random.Generator(np.random.PCG64(seed))
        x_act = rng.normal(0, 0.1, (B, n))
        x, P = np.zeros((B, n)), np.tile(P0, (B, 1, 1))
        out = dict(Q=Q, R=R, P0=P0, u=np.array(u), dt=np.array(dt), sqrt_mode=np.array(sqrt_mode), order=np.array(order))
        sd = np.sqrt(np.diag(R))
        for t in range(T):
            z = np.empty((B, o))
            for b in range(B):
                x_act[b] = O.fx(oid, p, x_act[b], u, dt)
                z[b] = O.hx(oid, p, x_act[b]) + sd * rng.standard_normal(o)
            x, P, st = O.ukf_step_batch(oid, p, x, P, Q, R, u, z, dt, sqrt_mode, order)
            assert not st.any()
            out[f"z_{t}"], out[f"x_{t}"], out[f"P_{t}"] = z, x.copy(), P.copy()
        return out
    return run


CASES = {
    "mppi_L": _mppi(O.MODEL_L, 8, 0.1, 0.5, 3.0, (-20.0, 20.0)),
    "mppi_NL": _mppi(O.MODEL_NL, 100, 0.008, 0.5, 3.0, (-20.0, 20.0)),
    "mppi_NL6": _mppi(O.MODEL_NL6, 8, 0.15, 1.4, 4.0, (-10.0, 10.0)),
    "ukf_PEN_LIN": _ukf(O.MODEL_PEN_LIN, O.SQRT_CHOLESKY, O.ORDER_INTERLEAVED, 0.0015),
    "ukf_PEN_NL": _ukf(O.MODEL_PEN_NL, O.SQRT_EIG, O.ORDER_LIBRARY, 0.1),
    "ukf_PEN6": _ukf(O.MODEL_PEN6, O.SQRT_EIG, O.ORDER_LIBRARY, 0.1),
    "ukf_NL6_UKF": _ukf(O.MODEL_NL6_UKF, O.SQRT_EIG, O.ORDER_LIBRARY, 0.3, dt=0.01),
}

if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    for name, fn in CASES.items():
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **fn())
        print("wrote", name)
