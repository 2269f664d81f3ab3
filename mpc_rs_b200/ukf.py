"""Host-side mirror of mpc::ukf::UnscentedKalmanFilter (n=4,o=3, src/ukf.rs), mpc::ukf2::UnscentedKalmanFilter
(n=6,o=5, src/ukf2.rs) and the free-function UKF of examples/ukf-pen.rs (n=4,o=2), batched over B filters.

    reference (Rust)                                   here
    let mut ukf = UnscentedKalmanFilter::new(x,p,q,r)   ukf = ukf.UnscentedKalmanFilter.new(x, p, q, r, fx=models.PEN_NL)
    ukf.predict(u, fx)                                  ukf.predict(u, models.PEN_NL)
    ukf.update(&z, hx)                                  ukf.update(z, models.PEN_NL)
    ukf.state() / ukf.covariance()                      ukf.state() / ukf.covariance()
    ukf.set_q(q)   (ukf2 only, src/ukf2.rs:96)          ukf.set_q(q);  ukf.set_r(r) (called by examples/mppi4-ukf-commu.rs:280)
    .expect("Inverse fail") panic                       UkfError("Inverse fail")

fx / hx are DeviceModel tags (the reference passes closures, src/ukf.rs:44-46,54-56).  With batch == 1 the
methods take and return the same shapes as the reference; with batch > 1 a leading [B] axis is added.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi as A
from .mppi import DeviceModel


class UkfError(RuntimeError):
    """The reference panics with this message (src/ukf.rs:69, examples/ukf-pen.rs:45)."""

    def __init__(self, status: int):
        super().__init__(A.status_string(status))
        self.status = status


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


class UserUkfModel(DeviceModel):
    """`fx` / `hx` of predict(u, fx) / update(&z, hx) (src/ukf.rs:44-46,54-56) as CUDA C++ source (mpcb_ukf_create_user):

        void fx(double (&x)[N], double u, double dt, const double* p);     // x <- f(x, u), in place
        void hx(const double (&x)[N], double (&z)[O], const double* p);    // z <- h(x)

    with N = n (1..6) and O = o (1..5) written as literals.  Build one with `user_ukf_model(source, n, o, params)`."""

    def __init__(self, source: str, n: int, o: int, params=(), name: str = "user-ukf"):
        object.__setattr__(self, "model_id", A.MODEL_USER_UKF)
        object.__setattr__(self, "name", name)
        object.__setattr__(self, "source", source)
        object.__setattr__(self, "n", int(n))
        object.__setattr__(self, "o", int(o))
        object.__setattr__(self, "params", tuple(float(v) for v in params))


def user_ukf_model(source: str, n: int, o: int, params=(), name: str = "user-ukf") -> UserUkfModel:
    if len(params) > A.USER_PARAMS:
        raise ValueError(f"at most {A.USER_PARAMS} parameters")
    return UserUkfModel(source, n, o, params, name)


def check_user_ukf_source(source: str, n: int, o: int) -> str:
    """Compile-only check (NVRTC, no GPU needed); returns the compiler log, raises MpcB200Error if it does not compile."""
    st = A.lib().mpcb_ukf_check_user_source(source.encode(), int(n), int(o))
    log = (A.lib().mpcb_rtc_log() or b"").decode()
    if st != A.OK:
        raise A.MpcB200Error(st, "user model did not compile:\n" + log)
    return log


_DIMS = {A.MODEL_PEN_LIN: (4, 2), A.MODEL_PEN_NL: (4, 3), A.MODEL_PEN6: (6, 5), A.MODEL_NL6_UKF: (6, 5)}


def default_noise(model: DeviceModel, dt: float = 0.0):
    """(Q, R, P0) the model's example ships with."""
    n, o = _DIMS[model.model_id]
    Q, R, P0 = np.empty((n, n)), np.empty((o, o)), np.empty((n, n))
    A.check(A.lib().mpcb_ukf_default_noise(model.model_id, float(dt), _dp(Q), _dp(R), _dp(P0)))
    return Q, R, P0


class BatchedUkf:
    """B independent unscented Kalman filters on one GPU (FP64, structure-of-arrays on the device)."""

    def __init__(self, model: DeviceModel, batch: int = 1, *, sqrt_mode: str = None, sigma_order: str = None,
                 device: int = 0, dt: float = None, params: dict = None, exact: bool = False):
        L = A.lib()
        cfg = A.UkfCfg()
        A.check(L.mpcb_ukf_default_cfg(model.model_id, C.byref(cfg)))
        user = isinstance(model, UserUkfModel)
        if user:
            cfg.n, cfg.o = model.n, model.o
        cfg.batch, cfg.device, cfg.exact = int(batch), int(device), int(bool(exact))
        if sqrt_mode is not None:
            cfg.sqrt_mode = {"cholesky": A.SQRT_CHOLESKY, "eig": A.SQRT_EIG, "svd": A.SQRT_EIG}[sqrt_mode]
        if sigma_order is not None:
            cfg.sigma_order = {"library": A.ORDER_LIBRARY, "interleaved": A.ORDER_INTERLEAVED}[sigma_order]
        if dt is not None:
            cfg.model.dt = float(dt)
        for k, v in ({} if user else (params or {})).items():
            setattr(cfg.model, k, float(v))
        self.cfg, self.model = cfg, model
        self.n, self.o, self.B = cfg.n, cfg.o, int(batch)
        self._h = A._H()
        if user:
            pa = (C.c_double * max(1, len(model.params)))(*model.params)
            st = L.mpcb_ukf_create_user(C.byref(self._h), C.byref(cfg), model.source.encode(), pa, len(model.params))
            if st == A.RTC_ERROR:
                raise A.MpcB200Error(st, "user model did not compile: " + L.mpcb_last_error_string().decode() + "\n"
                                      + (L.mpcb_rtc_log() or b"").decode())
            A.check(st)
        else:
            A.check(L.mpcb_ukf_create(C.byref(self._h), C.byref(cfg)))

    # -- lifetime --
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value is not None:
            A.lib().mpcb_ukf_destroy(self._h)
            self._h = A._H()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- state --
    def init(self, x, P, Q, R):
        """UnscentedKalmanFilter::new(x, p, q, r) for every filter of the batch."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(self.n)
        P = np.ascontiguousarray(P, dtype=np.float64).reshape(self.n, self.n)
        Q = np.ascontiguousarray(Q, dtype=np.float64).reshape(self.n, self.n)
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(self.o, self.o)
        A.check(A.lib().mpcb_ukf_init(self._h, _dp(x), _dp(P), _dp(Q), _dp(R)))

    def set_state(self, x=None, P=None):
        if x is not None:
            x = np.ascontiguousarray(x, dtype=np.float64).reshape(self.B, self.n)
        if P is not None:
            P = np.ascontiguousarray(P, dtype=np.float64).reshape(self.B, self.n, self.n)
        A.check(A.lib().mpcb_ukf_set_state(self._h, _dp(x), _dp(P)))

    def get_state(self, first: int = 0, count: int = None):
        count = self.B - first if count is None else count
        x, P = np.empty((count, self.n)), np.empty((count, self.n, self.n))
        A.check(A.lib().mpcb_ukf_get_state_range(self._h, first, count, _dp(x), _dp(P)))
        return x, P

    def set_q(self, Q):
        Q = np.ascontiguousarray(Q, dtype=np.float64).reshape(self.n, self.n)
        A.check(A.lib().mpcb_ukf_set_q(self._h, _dp(Q)))

    def set_r(self, R):
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(self.o, self.o)
        A.check(A.lib().mpcb_ukf_set_r(self._h, _dp(R)))

    def set_enable(self, enable: int):
        """Sensor bit mask of the following updates: a cleared bit zeroes that row of hx, like the closure of
        examples/mppi4-ukf-commu.rs:279-293.  Pair it with set_r(gen_r(enable, R))."""
        A.check(A.lib().mpcb_ukf_set_enable(self._h, int(enable) & 0xFFFFFFFF))

    def gen_r(self, enable: int, R):
        """gen_r of examples/mppi4-ukf-commu.rs:228-236: R with the variance of every disabled sensor set to 1e6."""
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(self.o, self.o)
        out = np.empty_like(R)
        A.check(A.lib().mpcb_ukf_gen_r(self._h, int(enable) & 0xFFFFFFFF, _dp(R), _dp(out)))
        return out

    def gather_state_device(self, idx, d_out: int):
        """Asynchronous: d_out[B][len(idx)] <- the state components idx of every filter (for Mppi.compute_device)."""
        arr = (C.c_int32 * len(idx))(*[int(i) for i in idx])
        A.check(A.lib().mpcb_ukf_gather_state_device(self._h, len(idx), arr, d_out))

    # -- filtering --
    def _u(self, u):
        if np.ndim(u) == 0:
            return None, float(u)
        ua = np.ascontiguousarray(u, dtype=np.float64).reshape(self.B)
        return ua, 0.0

    def _raise_on_failure(self):
        s = np.zeros(self.B, dtype=np.int32)
        st = A.lib().mpcb_ukf_get_status(self._h, s.ctypes.data_as(C.POINTER(C.c_int32)))
        if st in (A.INVERSE_FAIL, A.CHOLESKY_FAIL):
            raise UkfError(st)
        A.check(st)

    def predict(self, u, dt: float = 0.0, check: bool = True):
        ua, us = self._u(u)
        A.check(A.lib().mpcb_ukf_predict(self._h, _dp(ua), us, float(dt)))
        if check:
            self._raise_on_failure()

    def update(self, z, check: bool = True):
        z = np.ascontiguousarray(z, dtype=np.float64).reshape(self.B, self.o)
        A.check(A.lib().mpcb_ukf_update(self._h, _dp(z)))
        if check:
            self._raise_on_failure()

    def step(self, u, z, dt: float = 0.0, check: bool = True):
        """Fused predict + update (one kernel, sigma points stay in registers)."""
        ua, us = self._u(u)
        z = np.ascontiguousarray(z, dtype=np.float64).reshape(self.B, self.o)
        A.check(A.lib().mpcb_ukf_step(self._h, _dp(ua), us, float(dt), _dp(z)))
        if check:
            self._raise_on_failure()

    def run_device(self, steps: int, d_z: int, u=0.0, d_u: int = 0, dt: float = 0.0):
        """Asynchronous: `steps` fused steps on device-resident z[steps][o][B] (and u[steps][B] if d_u)."""
        A.check(A.lib().mpcb_ukf_run_device(self._h, int(steps), d_u or None, float(u), float(dt), d_z))

    def sync(self):
        A.check(A.lib().mpcb_ukf_sync(self._h))

    def status(self):
        s = np.zeros(self.B, dtype=np.int32)
        A.lib().mpcb_ukf_get_status(self._h, s.ctypes.data_as(C.POINTER(C.c_int32)))
        return s

    @property
    def launches(self) -> int:
        return A.lib().mpcb_ukf_launches(self._h)

    @property
    def stream(self) -> int:
        return A.lib().mpcb_ukf_stream(self._h) or 0


    @property
    def device_x(self) -> int:
        """Raw device pointer of the state, structure of arrays x[n][B] (mpcb_ukf_device_x)."""
        return A.lib().mpcb_ukf_device_x(self._h) or 0

    @property
    def device_p(self) -> int:
        """Raw device pointer of the covariance, P[n*n][B] (mpcb_ukf_device_p)."""
        return A.lib().mpcb_ukf_device_p(self._h) or 0


class UnscentedKalmanFilter(BatchedUkf):
    """Single filter with the reference's method names and shapes (src/ukf.rs:30-94)."""

    @classmethod
    def new(cls, x, p, q, r, *, fx: DeviceModel, **kw) -> "UnscentedKalmanFilter":
        f = cls(fx, 1, **kw)
        f.init(x, p, q, r)
        return f

    def predict(self, u, fx: DeviceModel = None, dt: float = 0.0):  # noqa: D401 - mirrors predict(u, fx)
        if fx is not None and fx.model_id != self.model.model_id:
            raise ValueError("fx must be the device model the filter was built with")
        super().predict(float(u), dt)

    def update(self, z, hx: DeviceModel = None):
        if hx is not None and hx.model_id != self.model.model_id:
            raise ValueError("hx must be the device model the filter was built with")
        super().update(np.asarray(z, dtype=np.float64).reshape(1, self.o))

    def state(self):
        return self.get_state()[0][0]

    def covariance(self):
        return self.get_state()[1][0]
