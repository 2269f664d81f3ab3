"""BASELINE.json's full sizes, checked through properties that do not need the (slow) oracle at that size:

  config #2 / #5  MPPI K = 65 536 x H = 100 and K = 2^24 x H = 200: determinism (same seed, same call index -> the same
                  bits), n_finite = K, argmin inside the range, controls inside the limits, sharding invariance (G
                  handles over contiguous sample ranges == one handle: the Philox counter is the global sample index),
                  and the oracle on a subsample-sized twin of the same configuration (K = 8192) for the arithmetic.
  config #3       UKF B = 2^20 filters x 100 steps: every filter given identical inputs ends bit-identical (the batch
                  index must not leak into the arithmetic), a random subset of filters driven by distinct measurements
                  matches the oracle run on just that subset, no failure statuses.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from mpc_rs_b200 import BatchedUkf, Mppi, models, ukf
from mpc_rs_b200 import _abi as A

pytestmark = pytest.mark.gpu
X0 = np.array([0.5, 0.0, 0.1, 0.0])


def _dev(nbytes):
    p = C.c_void_p()
    A.check(A.lib().mpcb_device_alloc(0, nbytes, C.byref(p)))
    return p


@pytest.mark.parametrize("K,H", [(65536, 100), (1 << 24, 200)])
def test_mppi_full_size_properties(gpu_required, K, H):
    dt, lam, sig, lim = 0.8 / H, 0.5, 3.0, (-20.0, 20.0)
    u_n = np.linspace(-1.0, 1.0, H)
    kw = dict(model=models.NL, lam=lam, std_dev=sig, limit=lim, precision="f32", dt=dt, seed=20240001)
    with Mppi(H, K, **kw) as a, Mppi(H, K, **kw) as b:
        ua, ub = a.compute(X0, u_n), b.compute(X0, u_n)
        ia = a.last_call_info()[0]
        assert np.array_equal(ua, ub), "same seed and call index must give the same bits"
        assert ia["status"] == 0 and ia["n_finite"] == K and 0 <= ia["argmax"] < K
        assert np.all(np.isfinite(ua)) and np.all(ua >= lim[0]) and np.all(ua <= lim[1])
        assert ia["sum"] >= 1.0  # the best sample has weight exp(0)
        ua2 = a.compute(X0, u_n)  # next call index: new noise
        assert not np.array_equal(ua, ua2)
    # sharding invariance at full size: 4 contiguous shards on one device, partial rows combined
    G = 4
    hs = [Mppi(H, K, rank=r, world_size=G, **kw) for r in range(G)]
    PL = hs[0].partial_len
    rows = _dev(8 * PL * G)
    for r, h in enumerate(hs):
        h.compute_partial(X0, u_n, rows.value + 8 * PL * r)
    us = hs[0].combine(rows.value, G)
    assert hs[0].info[0]["argmax"] == ia["argmax"] and hs[0].info[0]["n_finite"] == K
    assert np.linalg.norm(us - ua) / np.linalg.norm(ua) < 2e-6
    for h in hs:
        h.close()
    A.lib().mpcb_device_free(0, rows)
    # the arithmetic of this configuration against the oracle, at a size the oracle does in seconds
    Ks = 8192
    rng = np.random.Generator(np.random.PCG64(5))
    eps = (sig * rng.standard_normal((Ks, H))).astype(np.float32)
    p = O.model_defaults(O.MODEL_NL, dt=dt)
    st, u_o, io, _ = O.mppi_compute(O.MODEL_NL, p, Ks, H, lam, sig, lim[0], lim[1], X0, u_n, eps.astype(np.float64))
    with Mppi(H, Ks, **kw) as m:
        u_g = m.compute_replay(X0, u_n, eps)
        assert m.info[0]["argmax"] == io["argmax"]
        assert np.linalg.norm(u_g - u_o) / np.linalg.norm(u_o) < 1e-5


def test_ukf_full_size_properties(gpu_required):
    B, T = 1 << 20, 100
    Q, R, P0 = ukf.default_noise(models.PEN_LIN)
    rng = np.random.default_rng(20240003)
    pick = np.sort(rng.choice(B, 64, replace=False))
    z_common = rng.normal(0, 0.7, (T, 2))          # every filter sees this ...
    z_pick = rng.normal(0, 0.7, (T, len(pick), 2))  # ... except the picked ones
    # device-resident z[T][o][B] built on the host in two halves to keep the staging small
    z = np.empty((T, 2, B))
    z[:] = z_common[:, :, None]
    z[:, :, pick] = np.transpose(z_pick, (0, 2, 1))
    d_z = _dev(z.nbytes)
    A.check(A.lib().mpcb_device_upload(0, d_z, z.ctypes.data_as(C.c_void_p), z.nbytes))
    with BatchedUkf(models.PEN_LIN, B) as f:
        f.init(np.zeros(4), P0, Q, R)
        f.run_device(T, d_z.value, u=0.0015)
        f.sync()
        assert not f.status().any()
        x, P = f.get_state()
    A.lib().mpcb_device_free(0, d_z)
    others = np.ones(B, dtype=bool)
    others[pick] = False
    xo, Po = x[others], P[others]
    assert np.all(xo == xo[0]) and np.all(Po == Po[0]), "identical inputs must give identical filters"
    # the picked filters against the oracle on just those 64
    p = O.model_defaults(O.MODEL_PEN_LIN)
    xr, Pr = np.zeros((len(pick), 4)), np.tile(P0, (len(pick), 1, 1))
    for t in range(T):
        xr, Pr, st = O.ukf_step_batch(O.MODEL_PEN_LIN, p, xr, Pr, Q, R, 0.0015, z_pick[t], 0.0, O.SQRT_CHOLESKY, O.ORDER_INTERLEAVED)
        assert not st.any()
    assert np.linalg.norm(x[pick] - xr) / np.linalg.norm(xr) < 1e-5
    assert np.linalg.norm(P[pick] - Pr) / np.linalg.norm(Pr) < 1e-5
