"""ctypes binding of oracle/libmpc_oracle.so — the CPU restatement of the reference algorithm.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (mpc_rs_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "libmpc_oracle.so")

MODEL_L, MODEL_NL, MODEL_NL6 = 0, 1, 2
MODEL_PEN_LIN, MODEL_PEN_NL, MODEL_PEN6, MODEL_NL6_UKF = 16, 17, 18, 19
SQRT_CHOLESKY, SQRT_EIG = 0, 1
ORDER_LIBRARY, ORDER_INTERLEAVED = 0, 1


class ModelParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("m1", "r_w", "m2", "l", "j1", "j2", "g", "kt", "dt")] + [
        ("cost", C.c_double * 12)
    ]


class MppiOut(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("argmax", C.c_int64),
        ("max", C.c_double),
        ("sum", C.c_double),
        ("n_finite", C.c_int64),
    ]


class Gaussian(C.Structure):
    _fields_ = [("mean", C.c_double), ("var", C.c_double)]


def build(force: bool = False) -> str:
    src = [os.path.join(_ROOT, "oracle", f) for f in ("mpc_oracle.c", "mpc_oracle.h", "models_impl.inc")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", os.path.join(_ROOT, "oracle")], check=True, capture_output=True)
    return _SO


_lib = None
_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)


def _ptr(a, ty=_dp):
    return None if a is None else a.ctypes.data_as(ty)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    PP = C.POINTER(ModelParams)
    L.orc_model_defaults.argtypes = [C.c_int, PP]
    L.orc_dynamics.argtypes = [C.c_int, PP, _dp, C.c_double, _dp]
    L.orc_cost.argtypes = [C.c_int, PP, _dp]
    L.orc_cost.restype = C.c_double
    L.orc_ddot.argtypes = [PP, _dp, C.c_double, C.c_double, _dp, _dp]
    L.orc_dynamics_short.argtypes = [PP, _dp, C.c_double, C.c_double, C.c_double, _dp]
    L.orc_fx.argtypes = [C.c_int, PP, _dp, C.c_double, C.c_double, _dp]
    L.orc_hx.argtypes = [C.c_int, PP, _dp, _dp]
    L.orc_gen_q.argtypes = [C.c_double, _dp]
    L.orc_ukf_default_noise.argtypes = [C.c_int, C.c_double, _dp, _dp, _dp]
    mp = [C.c_int, PP, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _dp, _dp]
    L.orc_mppi_compute.argtypes = mp + [_dp, _dp, _dp, C.POINTER(MppiOut)]
    L.orc_mppi_compute_f32.argtypes = mp + [_fp, _dp, _dp, C.POINTER(MppiOut)]
    L.orc_mppi_compute_cpu.argtypes = mp + [C.c_uint64, C.c_int, _dp, C.POINTER(MppiOut)]
    L.orc_normal_fill.argtypes = [C.c_uint64, _dp, C.c_int64]
    L.orc_ukf_weights.argtypes = [C.c_int, _dp, _dp]
    L.orc_ukf_sigma_points.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]
    L.orc_cholesky_lower.argtypes = [C.c_int, _dp, _dp]
    L.orc_sym_eig_sqrt.argtypes = [C.c_int, _dp, _dp]
    L.orc_inverse.argtypes = [C.c_int, _dp, _dp]
    L.orc_ukf_predict.argtypes = [C.c_int, PP, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, _dp]
    L.orc_ukf_update.argtypes = [C.c_int, PP, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]
    L.orc_ukf_step_batch.argtypes = [C.c_int, PP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, _dp, _dp, _dp, _dp,
                                     _dp, C.c_double, C.c_double, _dp, C.POINTER(C.c_int32), C.c_int]
    L.orc_ukf_step_batch_masked.argtypes = L.orc_ukf_step_batch.argtypes + [C.c_uint32]
    L.orc_gen_r.argtypes = [C.c_int, _dp, C.c_uint32, _dp]
    L.orc_gen_r.restype = None
    for f in ("add", "sub", "mul"):
        fn = getattr(L, "orc_gaussian_" + f)
        fn.argtypes = [Gaussian, Gaussian]
        fn.restype = Gaussian
    L.orc_gaussian_scale.argtypes = [Gaussian, C.c_double]
    L.orc_gaussian_scale.restype = Gaussian
    _lib = L
    return L


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def model_defaults(model_id: int, **over) -> ModelParams:
    p = ModelParams()
    assert lib().orc_model_defaults(model_id, C.byref(p)) == 0
    for k, v in over.items():
        setattr(p, k, v)
    return p


def dims(model_id: int):
    return {MODEL_PEN_LIN: (4, 2), MODEL_PEN_NL: (4, 3), MODEL_PEN6: (6, 5), MODEL_NL6_UKF: (6, 5)}[model_id]


def dynamics(model_id, p, x, u):
    x = _f64(x)
    out = np.empty(4)
    lib().orc_dynamics(model_id, C.byref(p), _ptr(x), float(u), _ptr(out))
    return out


def cost(model_id, p, x):
    x = _f64(x)
    return lib().orc_cost(model_id, C.byref(p), _ptr(x))


def dynamics_short(p, x6, u, dt, f):
    x6 = _f64(x6)
    out = np.empty(6)
    lib().orc_dynamics_short(C.byref(p), _ptr(x6), float(u), float(dt), float(f), _ptr(out))
    return out


def fx(model_id, p, x, u, dt=0.0):
    x = _f64(x)
    out = np.empty(x.shape[0])
    lib().orc_fx(model_id, C.byref(p), _ptr(x), float(u), float(dt), _ptr(out))
    return out


def hx(model_id, p, x):
    x = _f64(x)
    out = np.empty(dims(model_id)[1])
    lib().orc_hx(model_id, C.byref(p), _ptr(x), _ptr(out))
    return out


def gen_q(dt):
    q = np.empty((6, 6))
    lib().orc_gen_q(float(dt), _ptr(q))
    return q


def ukf_default_noise(model_id, dt=0.0):
    n, o = dims(model_id)
    Q, R, P0 = np.empty((n, n)), np.empty((o, o)), np.empty((n, n))
    assert lib().orc_ukf_default_noise(model_id, float(dt), _ptr(Q), _ptr(R), _ptr(P0)) == 0
    return Q, R, P0


def mppi_compute(model_id, p, K, H, lam, std_dev, lo, hi, x, u_n, eps, f32=False, want_costs=False):
    """Returns (status, u_out[H], info dict, c[K] or None)."""
    x, u_n = _f64(x), _f64(u_n)
    u_out = np.zeros(H)
    c = np.empty(K) if want_costs else None
    info = MppiOut()
    if f32:
        eps = np.ascontiguousarray(eps, dtype=np.float32)
        assert eps.shape == (K, H)
        st = lib().orc_mppi_compute_f32(model_id, C.byref(p), K, H, lam, std_dev, lo, hi, _ptr(x), _ptr(u_n),
                                        _ptr(eps, _fp), _ptr(u_out), _ptr(c), C.byref(info))
    else:
        eps = _f64(eps)
        assert eps.shape == (K, H)
        st = lib().orc_mppi_compute(model_id, C.byref(p), K, H, lam, std_dev, lo, hi, _ptr(x), _ptr(u_n), _ptr(eps),
                                    _ptr(u_out), _ptr(c), C.byref(info))
    d = dict(status=info.status, argmax=info.argmax, max=info.max, sum=info.sum, n_finite=info.n_finite)
    return st, u_out, d, c


def mppi_compute_cpu(model_id, p, K, H, lam, std_dev, lo, hi, x, u_n, seed=1, threads=0):
    x, u_n = _f64(x), _f64(u_n)
    u_out = np.zeros(H)
    info = MppiOut()
    st = lib().orc_mppi_compute_cpu(model_id, C.byref(p), K, H, lam, std_dev, lo, hi, _ptr(x), _ptr(u_n), seed,
                                    threads, _ptr(u_out), C.byref(info))
    return st, u_out, dict(status=info.status, argmax=info.argmax, max=info.max, sum=info.sum, n_finite=info.n_finite)


def max_threads():
    return lib().orc_max_threads()


def normal_fill(seed, n):
    out = np.empty(n)
    lib().orc_normal_fill(seed, _ptr(out), n)
    return out


def ukf_weights(n):
    wm, wc = np.empty(2 * n + 1), np.empty(2 * n + 1)
    lib().orc_ukf_weights(n, _ptr(wm), _ptr(wc))
    return wm, wc


def cholesky_lower(A):
    A = _f64(A)
    n = A.shape[0]
    L = np.empty((n, n))
    st = lib().orc_cholesky_lower(n, _ptr(A), _ptr(L))
    return st, L


def sym_eig_sqrt(A):
    A = _f64(A)
    n = A.shape[0]
    L = np.empty((n, n))
    lib().orc_sym_eig_sqrt(n, _ptr(A), _ptr(L))
    return L


def inverse(A):
    A = _f64(A)
    n = A.shape[0]
    Ai = np.empty((n, n))
    st = lib().orc_inverse(n, _ptr(A), _ptr(Ai))
    return st, Ai


def ukf_sigma_points(x, P, sqrt_mode, order):
    x, P = _f64(x), _f64(P)
    n = x.shape[0]
    sig = np.empty((n, 2 * n + 1))
    st = lib().orc_ukf_sigma_points(n, sqrt_mode, order, _ptr(x), _ptr(P), _ptr(sig))
    return st, sig


def ukf_predict(model_id, p, x, P, Q, u, dt=0.0, sqrt_mode=SQRT_CHOLESKY, order=ORDER_LIBRARY):
    """Returns (status, x', P', sigma_f[n][M])."""
    x, P, Q = _f64(x).copy(), _f64(P).copy(), _f64(Q)
    n = x.shape[0]
    sf = np.empty((n, 2 * n + 1))
    st = lib().orc_ukf_predict(model_id, C.byref(p), n, sqrt_mode, order, _ptr(x), _ptr(P), _ptr(Q), float(u),
                               float(dt), _ptr(sf))
    return st, x, P, sf


def ukf_update(model_id, p, x, P, R, z, sigma_f):
    x, P, R, z, sigma_f = _f64(x).copy(), _f64(P).copy(), _f64(R), _f64(z), _f64(sigma_f)
    n, o = x.shape[0], z.shape[0]
    st = lib().orc_ukf_update(model_id, C.byref(p), n, o, _ptr(x), _ptr(P), _ptr(R), _ptr(z), _ptr(sigma_f))
    return st, x, P


def gen_r(R, enable):
    """gen_r of examples/mppi4-ukf-commu.rs:228-236."""
    R = _f64(R)
    o = R.shape[0]
    out = np.empty_like(R)
    lib().orc_gen_r(o, _ptr(R), enable, _ptr(out))
    return out


def ukf_step_batch(model_id, p, x, P, Q, R, u, z, dt=0.0, sqrt_mode=SQRT_CHOLESKY, order=ORDER_LIBRARY, threads=0,
                   enable=0xFFFFFFFF):
    """x[B][n], P[B][n][n] AoS (copied), u scalar or [B], z[B][o]. Returns (x', P', status[B]).
    enable: sensor bit mask of examples/mppi4-ukf-commu.rs:279-293 (hx rows of cleared bits are zeroed)."""
    x, P = _f64(x).copy(), _f64(P).copy()
    B, n = x.shape
    z = _f64(z)
    o = z.shape[1]
    status = np.zeros(B, dtype=np.int32)
    if np.ndim(u) == 0:
        up, us = None, float(u)
    else:
        ua = _f64(u)
        up, us = _ptr(ua), 0.0
    Q, R = _f64(Q), _f64(R)
    lib().orc_ukf_step_batch_masked(model_id, C.byref(p), n, o, sqrt_mode, order, B, _ptr(x), _ptr(P), _ptr(Q), _ptr(R), up,
                                    us, float(dt), _ptr(z), status.ctypes.data_as(C.POINTER(C.c_int32)), threads, enable)
    return x, P, status
