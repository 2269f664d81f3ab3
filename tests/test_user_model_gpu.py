"""User-supplied MPPI models (mpcb_mppi_create_user): the reference's `dynamics` / `cost` fn pointers of Mppi::new
(src/mppi.rs:9-10,16-22) handed over as CUDA source and compiled into the fused kernel at create time.

Parity: the reference's own nonlinear model written as a user model must give what the oracle gives for model NL
(and therefore what the built-in kernel gives); a model the library has never seen is checked against the independent
numpy restatement of src/mppi.rs (tests/ref_numpy.py) with the same dynamics/cost written in numpy."""
import numpy as np
import pytest

import oracle_lib as O
import ref_numpy as R
from mpc_rs_b200 import Mppi, MppiError, MpcB200Error, models, user_model
from mpc_rs_b200 import _abi as A

pytestmark = pytest.mark.gpu

X0 = np.array([0.5, 0.0, 0.1, 0.0])

# examples/mppi4-non-liner.rs:20-27,73-94 as a user would port it: same association order, constants as parameters
NL_SOURCE = r"""
template <typename real>
void dynamics(real (&x)[4], real u, const real* p) {
    const real M1 = p[0], R_W = p[1], M2 = p[2], L = p[3], J1 = p[4], J2 = p[5], G = p[6], KT = p[7], DT = p[8];
    const real s = sin(x[2]), c = cos(x[2]);
    const real D = (M1 + M2 + J1 / (R_W * R_W)) * (M2 * L * L + J2);
    const real d = D - M2 * M2 * L * L * c * c;
    const real term1 = (M1 + M2 + J1 / (R_W * R_W)) * M2 * G * L * s;
    const real drive = KT * u / R_W + M2 * L * (x[3] * x[3]) * s;
    const real term2 = drive * M2 * L * c;
    const real r3 = x[3] + (term1 - term2) / d * DT;
    const real r2 = x[2] + x[3] * DT;
    const real term3 = (J2 + M2 * L * L) * drive;
    const real term4 = M2 * G * L * L * s * c;
    const real r1 = x[1] + (term3 + term4) / d * DT;
    const real r0 = x[0] + x[1] * DT;
    x[0] = r0; x[1] = r1; x[2] = r2; x[3] = r3;
}
template <typename real>
real cost(const real (&x)[4], const real* p) {
    const real xc = mpcb::clampm(x[0], (real)-2.0, (real)2.0);
    const real a = mpcb::clampm(x[1] + (real)2.0 * xc, (real)-5.0, (real)5.0);
    const real b = x[2] + (real)0.35 * mpcb::clampm(x[0], (real)-0.75, (real)0.75);
    return (real)2.0 * (xc * xc) + (real)3.0 * (a * a) + (real)5.0 * (b * b) + (real)1.2 * (x[3] * x[3]);
}
"""


def rel_err(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def nl_params(dt):
    p = O.model_defaults(O.MODEL_NL, dt=dt)
    return [p.m1, p.r_w, p.m2, p.l, p.j1, p.j2, p.g, p.kt, dt]


@pytest.mark.parametrize("precision,tol", [("f64", 1e-9), ("f32", 1e-5)])
@pytest.mark.parametrize("H,dt,K", [(8, 0.1, 16384), (100, 0.008, 65536)])
def test_reference_model_as_user_source_matches_the_oracle(gpu_required, precision, tol, H, dt, K):
    lam, sig, lim = 0.5, 3.0, (-20.0, 20.0)
    p = O.model_defaults(O.MODEL_NL, dt=dt)
    rng = np.random.Generator(np.random.PCG64(11))
    um = user_model(NL_SOURCE, nl_params(dt), name="NL-as-user")
    x, u = X0.copy(), np.zeros(H)
    with Mppi(H, K, model=um, lam=lam, std_dev=sig, limit=lim, precision=precision) as m, \
            Mppi(H, K, model=models.NL, lam=lam, std_dev=sig, limit=lim, precision=precision, dt=dt) as builtin:
        assert m.cfg.model_id == A.MODEL_USER
        for _ in range(3):
            eps = (sig * rng.standard_normal((K, H))).astype(np.float64 if precision == "f64" else np.float32)
            st, u_o, io, _ = O.mppi_compute(O.MODEL_NL, p, K, H, lam, sig, lim[0], lim[1], x, u, eps.astype(np.float64))
            assert st == 0
            u_g = m.compute_replay(x, u, eps)
            u_b = builtin.compute_replay(x, u, eps)
            assert m.info[0]["argmax"] == io["argmax"] == builtin.info[0]["argmax"]
            assert m.info[0]["n_finite"] == io["n_finite"]
            assert rel_err(u_g, u_o) < tol, rel_err(u_g, u_o)
            assert rel_err(u_g, u_b) < 2 * tol
            x = O.dynamics(O.MODEL_NL, p, x, u_o[0])
            u = u_o.copy()


# a model the library has never seen: a cart with drag steering a spring-coupled second mass, tanh-saturated input
CUSTOM_SOURCE = r"""
template <typename real>
void dynamics(real (&x)[4], real u, const real* p) {
    const real dt = p[0], k = p[1], drag = p[2], gain = p[3];
    const real f = gain * tanh(u);
    const real a0 = f - drag * x[1] * fabs(x[1]) - k * (x[0] - x[2]);
    const real a1 = k * (x[0] - x[2]) - (real)0.5 * x[3];
    x[0] = x[0] + x[1] * dt;
    x[1] = x[1] + a0 * dt;
    x[2] = x[2] + x[3] * dt;
    x[3] = x[3] + a1 * dt;
}
template <typename real>
real cost(const real (&x)[4], const real* p) {
    const real e = x[2] - p[4];
    return e * e + (real)0.1 * (x[1] * x[1]) + (real)0.05 * (x[3] * x[3]) + (real)0.3 * fabs(x[0] - x[2]);
}
"""
CUSTOM_PARAMS = [0.05, 4.0, 0.3, 2.5, 1.0]


def custom_dyn(x, u, dt_unused):
    dt, k, drag, gain = CUSTOM_PARAMS[:4]
    x = np.asarray(x, dtype=np.float64)
    f = gain * np.tanh(u)
    a0 = f - drag * x[..., 1] * np.abs(x[..., 1]) - k * (x[..., 0] - x[..., 2])
    a1 = k * (x[..., 0] - x[..., 2]) - 0.5 * x[..., 3]
    return np.stack([x[..., 0] + x[..., 1] * dt, x[..., 1] + a0 * dt, x[..., 2] + x[..., 3] * dt, x[..., 3] + a1 * dt], axis=-1)


def custom_cost(x):
    e = x[..., 2] - CUSTOM_PARAMS[4]
    return e * e + 0.1 * x[..., 1] ** 2 + 0.05 * x[..., 3] ** 2 + 0.3 * np.abs(x[..., 0] - x[..., 2])


@pytest.mark.parametrize("precision,tol", [("f64", 1e-10), ("f32", 1e-5)])
def test_unseen_model_against_numpy_mppi(gpu_required, precision, tol):
    R.MPPI_MODELS["custom"] = (custom_dyn, custom_cost)
    lam, sig, lim = 0.8, 1.5, (-3.0, 3.0)
    rng = np.random.default_rng(3)
    um = user_model(CUSTOM_SOURCE, CUSTOM_PARAMS)
    for K, H, C in ((4096, 20, 1), (777, 7, 1), (2048, 12, 3)):
        with Mppi(H, K, model=um, lam=lam, std_dev=sig, limit=lim, precision=precision, controllers=C, keep_costs=True) as m:
            xs = rng.normal(0, 0.5, (C, 4))
            us = rng.uniform(-1, 1, (C, H))
            eps = (sig * rng.standard_normal((C, K, H))).astype(np.float64 if precision == "f64" else np.float32)
            u_g = np.reshape(m.compute_replay(xs, us, eps), (C, H))
            c_g = np.reshape(m.costs(), (C, K))
            for c in range(C):
                u_r, c_r = R.mppi_compute("custom", 0.0, lam, sig, lim, xs[c], us[c], eps[c].astype(np.float64))
                assert m.info[c]["argmax"] == int(np.argmax(c_r))
                assert rel_err(u_g[c], u_r) < tol, (K, H, c, rel_err(u_g[c], u_r))
                np.testing.assert_allclose(c_g[c], c_r, rtol=1e-9 if precision == "f64" else 2e-4, atol=1e-9 if precision == "f64" else 1e-3)


def test_user_model_generate_mode_and_sharding(gpu_required):
    """Generate mode draws the same Philox noise as the built-in path, so the user port of model NL tracks the built-in
    kernel there too; a sharded pair of user handles combines to the single-handle result (compute_partial + combine)."""
    H, K, dt = 16, 40000, 0.05
    lam, sig, lim = 0.5, 3.0, (-20.0, 20.0)
    um = user_model(NL_SOURCE, nl_params(dt))
    with Mppi(H, K, model=um, lam=lam, std_dev=sig, limit=lim, precision="f64", seed=9) as m, \
            Mppi(H, K, model=models.NL, lam=lam, std_dev=sig, limit=lim, precision="f64", dt=dt, seed=9) as b:
        u = np.zeros(H)
        for _ in range(3):
            u_m, u_b = m.compute(X0, u), b.compute(X0, u)
            assert m.last_call_info()[0]["argmax"] == b.last_call_info()[0]["argmax"]
            assert rel_err(u_m, u_b) < 1e-9
            u = u_b
        one = m.compute(X0, u)
    import ctypes as C
    G = 2
    hs = [Mppi(H, K, model=um, lam=lam, std_dev=sig, limit=lim, precision="f64", seed=9, rank=r, world_size=G) for r in range(G)]
    PL = hs[0].partial_len
    d = C.c_void_p()
    A.check(A.lib().mpcb_device_alloc(0, 8 * PL * G, C.byref(d)))
    # same call index as `one` (the 4th compute): burn three steps on each shard
    for h in hs:
        for _ in range(3):
            h.compute_partial(X0, np.zeros(H), d.value)
    for r, h in enumerate(hs):
        h.compute_partial(X0, u, d.value + 8 * PL * r)
        h.sync()
    comb = hs[0].combine(d.value, G)
    assert rel_err(comb, one) < 1e-12
    for h in hs:
        h.close()
    A.lib().mpcb_device_free(0, d)


def test_user_model_errors(gpu_required):
    with pytest.raises(MpcB200Error) as e:
        Mppi(8, 1024, model=user_model("void dynamics(float (&x)[4], float u, const float* p) { x[0] = undefined_name; }"),
             lam=1.0, std_dev=1.0)
    assert e.value.status == A.RTC_ERROR and "undefined_name" in str(e.value) and "user_model.cu(1)" in str(e.value)
    # a cost that is NaN everywhere: the reference's Err("Cannot calculate max") (src/mppi.rs:69)
    src = """
    void dynamics(float (&x)[4], float u, const float* p) { x[0] += u; }
    float cost(const float (&x)[4], const float* p) { return sqrtf(-1.0f - x[0] * x[0]); }
    """
    with Mppi(8, 1024, model=user_model(src), lam=1.0, std_dev=1.0, precision="f32") as m:
        with pytest.raises(MppiError, match="Cannot calculate max"):
            m.compute(X0, np.zeros(8))


def test_example_user_model_parks_the_cart(gpu_required):
    """examples/mppi4_user_model.py: a model the library does not ship, closed loop for 6 s: the cart ends at the 1 m
    target with the load at rest, inside the control limits."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
    import mppi4_user_model as ex
    rows = ex.run(samples=32768, seconds=6.0, quiet=True, seed=3)
    assert np.all(np.isfinite(rows)) and np.all(np.abs(rows[:, 1]) <= 3.0)
    assert abs(rows[-1, 2] - 1.0) < 0.15 and abs(rows[-1, 4]) < 0.1 and abs(rows[-1, 3]) < 0.3


# ---------------------------------------------------------------------------------------------------------------
# UKF: fx / hx closures of predict(u, fx) / update(&z, hx) (src/ukf.rs:44-46,54-56) as CUDA source
# ---------------------------------------------------------------------------------------------------------------
from mpc_rs_b200 import BatchedUkf, UnscentedKalmanFilter, user_ukf_model  # noqa: E402

# examples/ukf-pen2.rs:31-53 the way a user would port it (p = M1, R_W, M2, L, J1, J2, G, KT)
PEN_NL_SOURCE = r"""
void fx(double (&x)[4], double u, double dt, const double* p) {
    const double M1 = p[0], R_W = p[1], M2 = p[2], L = p[3], J1 = p[4], J2 = p[5], G = p[6], KT = p[7];
    const double s = sin(x[2]), c = cos(x[2]);
    const double D = (M1 + M2 + J1 / (R_W * R_W)) * (M2 * L * L + J2);
    const double d = D - M2 * M2 * L * L * c * c;
    const double term1 = (M1 + M2 + J1 / (R_W * R_W)) * M2 * G * L * s;
    const double drive = KT * u / R_W + M2 * L * (x[3] * x[3]) * s;
    const double term2 = drive * M2 * L * c;
    const double r3 = x[3] + (term1 - term2) / d * dt;
    const double r2 = x[2] + x[3] * dt;
    const double term3 = (J2 + M2 * L * L) * drive;
    const double term4 = M2 * G * L * L * s * c;
    const double r1 = x[1] + (term3 + term4) / d * dt;
    const double r0 = x[0] + x[1] * dt;
    x[0] = r0; x[1] = r1; x[2] = r2; x[3] = r3;
}
void hx(const double (&x)[4], double (&z)[3], const double* p) {
    const double PI = 3.14159265358979323846264338327950288;
    z[0] = 60.0 / (2.0 * PI * p[1]) * x[1];
    z[1] = 60.0 / (2.0 * PI * p[1]) * x[1];
    z[2] = x[3] * (180.0 / PI);
}
"""


# the port recomputes D, 60/(2 pi R_W), ... per call where the built-in model folds them on the host: rounding-level
# differences times the 1.7e5 sigma-weight amplification
@pytest.mark.parametrize("sqrt_mode,exact,tol", [("eig", True, 1e-8), ("cholesky", True, 1e-8), ("eig", False, 1e-7)])
def test_reference_ukf_model_as_user_source_matches_the_oracle(gpu_required, sqrt_mode, exact, tol):
    B, T, u, dt = 129, 5, 0.1, 0.01
    oid = O.MODEL_PEN_NL
    p = O.model_defaults(oid)
    Q, R, P0 = O.ukf_default_noise(oid, 0.0)
    rng = np.random.default_rng(8)
    um = user_ukf_model(PEN_NL_SOURCE, 4, 3, [p.m1, p.r_w, p.m2, p.l, p.j1, p.j2, p.g, p.kt])
    x = rng.normal(0, 0.1, (B, 4))
    P = np.tile(P0, (B, 1, 1))
    osq = {"eig": O.SQRT_EIG, "cholesky": O.SQRT_CHOLESKY}[sqrt_mode]
    with BatchedUkf(um, B, sqrt_mode=sqrt_mode, sigma_order="library", exact=exact, dt=dt) as f:
        assert (f.n, f.o) == (4, 3)
        f.init(np.zeros(4), P0, Q, R)
        for t in range(T):
            z = rng.normal(0, 1.0, (B, 3)) * np.sqrt(np.diag(R))
            f.set_state(x, P)
            if t % 2:
                f.predict(u, dt)
                f.update(z)
            else:
                f.step(u, z, dt)
            xg, Pg = f.get_state()
            x, P, st = O.ukf_step_batch(oid, p, x, P, Q, R, u, z, dt, osq, O.ORDER_LIBRARY)
            assert not st.any()
            assert rel_err(xg, x) < tol and rel_err(Pg, P) < tol, (t, rel_err(xg, x), rel_err(Pg, P))


# a filter the library has never seen: n = 3, o = 2 (damped pendulum with a slowly adapting bias; range + mixed sensor)
UKF3_SOURCE = r"""
void fx(double (&x)[3], double u, double dt, const double* p) {
    const double th = x[0], om = x[1], b = x[2];
    x[0] = th + om * dt;
    x[1] = om + (u - p[0] * sin(th) - p[1] * om + b) * dt;
    x[2] = 0.98 * b;
}
void hx(const double (&x)[3], double (&z)[2], const double* p) {
    z[0] = p[2] * sin(x[0]);
    z[1] = x[1] + 0.1 * x[0] * x[0];
}
"""
UKF3_PARAMS = [9.81 / 0.5, 0.2, 1.5]


def ukf3_fx(x, u, dt):
    a, bq, _ = UKF3_PARAMS
    return np.array([x[0] + x[1] * dt, x[1] + (u - a * np.sin(x[0]) - bq * x[1] + x[2]) * dt, 0.98 * x[2]])


def ukf3_hx(x):
    return np.array([UKF3_PARAMS[2] * np.sin(x[0]), x[1] + 0.1 * x[0] * x[0]])


@pytest.mark.parametrize("sqrt_mode,order", [("cholesky", "interleaved"), ("eig", "library")])
def test_unseen_ukf_model_against_numpy(gpu_required, sqrt_mode, order):
    dt, u, T = 0.02, 0.3, 6
    Q = np.diag([1e-4, 1e-2, 1e-3])
    Rm = np.diag([0.05, 0.2])
    P = np.array([[1.0, 0.1, 0.0], [0.1, 2.0, 0.2], [0.0, 0.2, 3.0]])  # distinct eigenvalues: the eigen square root is unique
    x = np.array([0.4, -0.2, 0.05])
    rng = np.random.default_rng(5)
    f = UnscentedKalmanFilter.new(x, P, Q, Rm, fx=user_ukf_model(UKF3_SOURCE, 3, 2, UKF3_PARAMS), sqrt_mode=sqrt_mode,
                                  sigma_order=order, exact=True)
    for t in range(T):
        z = ukf3_hx(x) + rng.normal(0, 0.1, 2)
        f.set_state(x[None, :], P[None, :, :])
        f.predict(u, dt=dt)
        f.update(z)
        xr, Pr, sig = R_ukf_predict(x, P, Q, u, dt, sqrt_mode, order)
        x, P = R.ukf_update(ukf3_hx, xr, Pr, Rm, z, sig)
        assert rel_err(f.state(), x) < 1e-8, (t, rel_err(f.state(), x))
        assert rel_err(f.covariance(), P) < 1e-8, (t, rel_err(f.covariance(), P))
    f.close()


def R_ukf_predict(x, P, Q, u, dt, sqrt_mode, order):
    return R.ukf_predict(lambda s: ukf3_fx(s, u, dt), x, P, Q, sqrt_mode="cholesky" if sqrt_mode == "cholesky" else "svd",
                         order=order)


def test_user_ukf_errors(gpu_required):
    with pytest.raises(MpcB200Error) as e:
        BatchedUkf(user_ukf_model("void fx(double (&x)[2], double u, double dt, const double* p) { x[0] = oops; }", 2, 1), 4)
    assert e.value.status == A.RTC_ERROR and "oops" in str(e.value)
    with pytest.raises(MpcB200Error):
        BatchedUkf(user_ukf_model("void fx(){}", 7, 1), 4)  # n out of range


# ---------------------------------------------------------------------------------------------------------------
# Mppi<N,K,S> is generic in S: user models with S = 2 and S = 6 against the numpy MPPI
# ---------------------------------------------------------------------------------------------------------------
S2_SOURCE = r"""
template <typename real> void dynamics(real (&x)[2], real u, const real* p) {
    const real v = x[1];
    x[0] += v * p[0];
    x[1] += (u - p[1] * v * fabs(v)) * p[0];
}
template <typename real> real cost(const real (&x)[2], const real* p) {
    const real e = x[0] - p[2];
    return e * e + (real)0.1 * (x[1] * x[1]);
}
"""
S2_PARAMS = [0.05, 0.4, 1.0]
S6_SOURCE = r"""
// three masses on a line coupled by springs, the control pushes the first: x = [q0, v0, q1, v1, q2, v2]
template <typename real> void dynamics(real (&x)[6], real u, const real* p) {
    const real dt = p[0], k = p[1], c = p[2];
    const real a0 = u - k * (x[0] - x[2]) - c * x[1];
    const real a1 = k * (x[0] - x[2]) - k * (x[2] - x[4]) - c * x[3];
    const real a2 = k * (x[2] - x[4]) - c * x[5] - (real)0.3 * sin(x[4]);
    x[0] += x[1] * dt; x[2] += x[3] * dt; x[4] += x[5] * dt;
    x[1] += a0 * dt; x[3] += a1 * dt; x[5] += a2 * dt;
}
template <typename real> real cost(const real (&x)[6], const real* p) {
    const real e = x[4] - p[3];
    return (real)2.0 * (e * e) + (real)0.05 * (x[1] * x[1] + x[3] * x[3] + x[5] * x[5]) + (real)0.2 * ((x[0] - x[4]) * (x[0] - x[4]));
}
"""
S6_PARAMS = [0.04, 6.0, 0.3, 0.5]


def _s2_dyn(x, u, _):
    dt, drag, _t = S2_PARAMS
    v = x[..., 1]
    return np.stack([x[..., 0] + v * dt, v + (u - drag * v * np.abs(v)) * dt], axis=-1)


def _s2_cost(x):
    e = x[..., 0] - S2_PARAMS[2]
    return e * e + 0.1 * x[..., 1] ** 2


def _s6_dyn(x, u, _):
    dt, k, c, _t = S6_PARAMS
    a0 = u - k * (x[..., 0] - x[..., 2]) - c * x[..., 1]
    a1 = k * (x[..., 0] - x[..., 2]) - k * (x[..., 2] - x[..., 4]) - c * x[..., 3]
    a2 = k * (x[..., 2] - x[..., 4]) - c * x[..., 5] - 0.3 * np.sin(x[..., 4])
    return np.stack([x[..., 0] + x[..., 1] * dt, x[..., 1] + a0 * dt, x[..., 2] + x[..., 3] * dt, x[..., 3] + a1 * dt,
                     x[..., 4] + x[..., 5] * dt, x[..., 5] + a2 * dt], axis=-1)


def _s6_cost(x):
    e = x[..., 4] - S6_PARAMS[3]
    return 2.0 * e * e + 0.05 * (x[..., 1] ** 2 + x[..., 3] ** 2 + x[..., 5] ** 2) + 0.2 * (x[..., 0] - x[..., 4]) ** 2


@pytest.mark.parametrize("S,source,params,dyn,cost", [(2, S2_SOURCE, S2_PARAMS, _s2_dyn, _s2_cost), (6, S6_SOURCE, S6_PARAMS, _s6_dyn, _s6_cost)])
@pytest.mark.parametrize("precision,tol", [("f64", 1e-10), ("f32", 1e-5)])
def test_other_state_dimensions(gpu_required, S, source, params, dyn, cost, precision, tol):
    R.MPPI_MODELS["custom_s"] = (dyn, cost)
    lam, sig, lim = 0.6, 1.2, (-4.0, 4.0)
    rng = np.random.default_rng(10 + S)
    um = user_model(source, params)
    # the long horizon (two merge levels, no v tile) only in FP64: 12 s of FP32 rollout is beyond the 1e-5 bar
    for K, H, C in ((4096, 16, 1), (1500, 9, 2), (40000, 300 if precision == "f64" else 40, 1)):
        with Mppi(H, K, S, model=um, lam=lam, std_dev=sig, limit=lim, precision=precision, controllers=C) as m:
            assert m.S == S
            xs = rng.normal(0, 0.4, (C, S))
            us = rng.uniform(-1, 1, (C, H))
            eps = (sig * rng.standard_normal((C, K, H))).astype(np.float64 if precision == "f64" else np.float32)
            u_g = np.reshape(m.compute_replay(xs if C > 1 else xs[0], us if C > 1 else us[0], eps if C > 1 else eps[0]), (C, H))
            for c in range(C):
                u_r, c_r = R.mppi_compute("custom_s", 0.0, lam, sig, lim, xs[c], us[c], eps[c].astype(np.float64))
                assert m.info[c]["argmax"] == int(np.argmax(c_r))
                assert rel_err(u_g[c], u_r) < tol, (S, K, H, c, rel_err(u_g[c], u_r))
            # generate mode on host buffers (x travels inline in the kernel parameters for one controller)
            ug = np.reshape(m.compute(xs if C > 1 else xs[0], us if C > 1 else us[0]), (C, H))
            assert np.all(np.isfinite(ug)) and np.all(np.abs(ug) <= 4.0)
    with pytest.raises(MpcB200Error):
        Mppi(8, 1024, 3, model=models.NL, lam=1.0, std_dev=1.0)  # the built-in models are S = 4


def test_example_user_ukf_finds_the_bias(gpu_required):
    """examples/ukf_user_model.py: n = 3, o = 2 with the caller's own fx / hx; 256 filters on noisy measurements of
    plants with an unknown torque bias of 1.5: every filter finds the bias and tracks angle and rate."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
    import ukf_user_model as ex
    x_act, x_est, p = ex.run(batch=256, steps=400, quiet=True, seed=1)
    assert np.all(np.isfinite(x_est)) and np.all(np.isfinite(p))
    assert np.max(np.abs(x_est[:, 2] - x_act[:, 2])) < 0.15      # bias found (prior 0 +- 2, truth 1.5)
    assert np.max(np.abs(x_est[:, 0] - x_act[:, 0])) < 0.05 and np.max(np.abs(x_est[:, 1] - x_act[:, 1])) < 0.1
    assert np.all(p[:, 2, 2] < 0.05)                             # and it knows it (P0 = 4)
