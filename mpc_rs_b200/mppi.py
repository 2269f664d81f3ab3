"""Host-side mirror of mpc::mppi::Mppi<N,K,S> (reference src/mppi.rs:7-92) over the C ABI.

    reference (Rust)                                             here
    Mppi::<N,K,4>::new(dynamics, cost, LAMBDA, R, LIMIT)         Mppi.new(models.L, models.L, LAMBDA, R, LIMIT, N=8, K=800_000)
    u_n = mppi.compute(&x, &u_n)?                                u_n = mppi.compute(x, u_n)     # raises MppiError
    Err("Cannot calculate max" | "sum is zero" | "u is invalid")  MppiError(str(e) is the same text)

The reference takes host fn pointers for dynamics/cost (src/mppi.rs:9-10); here they are DeviceModel tags
naming the built-in device models (mpc_rs_b200.models).  All arithmetic runs in libmpc_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _abi as A


@dataclass(frozen=True)
class DeviceModel:
    """Tag standing in for the reference's `dynamics` / `cost` fn pointers."""
    model_id: int
    name: str


@dataclass(frozen=True)
class UserModel(DeviceModel):
    """`dynamics` and `cost` of Mppi::new (src/mppi.rs:9-10) as CUDA C++ source + constants (mpcb_mppi_create_user):

        template <typename real> void dynamics(real (&x)[S], real u, const real* p);   // x <- f(x, u)
        template <typename real> real cost(const real (&x)[S], const real* p);

    S = the S of Mppi<N,K,S> (1..8, the `S` argument of Mppi).  Build one with `user_model(source, params)`."""
    source: str = ""
    params: tuple = ()


def user_model(source: str, params=(), name: str = "user") -> UserModel:
    params = tuple(float(v) for v in params)
    if len(params) > A.USER_PARAMS:
        raise ValueError(f"at most {A.USER_PARAMS} parameters")
    return UserModel(A.MODEL_USER, name, source, params)


def check_user_source(source: str, precision: str = "f32", S: int = 4) -> str:
    """Compile-only check (NVRTC, no GPU needed); returns the compiler log, raises MpcB200Error if it does not compile."""
    st = A.lib().mpcb_mppi_check_user_source(source.encode(), int(S), {"f32": A.F32, "f64": A.F64}[precision])
    log = (A.lib().mpcb_rtc_log() or b"").decode()
    if st != A.OK:
        raise A.MpcB200Error(st, "user model did not compile:\n" + log)
    return log


class MppiError(RuntimeError):
    """Err(&'static str) of Mppi::compute (src/mppi.rs:69,77,88); str(e) is the reference's message."""

    def __init__(self, status: int):
        super().__init__(A.status_string(status))
        self.status = status


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Mppi:
    """MPPI controller: N = horizon, K = samples, S = state dimension (src/mppi.rs:7)."""

    def __init__(self, N: int, K: int, S: int = 4, *, model: DeviceModel, lam: float, std_dev: float,
                 limit=(-float("inf"), float("inf")), precision: str = None, controllers: int = 1, seed: int = None,
                 device: int = 0, rank: int = 0, world_size: int = 1, keep_costs: bool = False, dt: float = None,
                 params: dict = None):
        L = A.lib()
        cfg = A.MppiCfg()
        A.check(L.mpcb_mppi_default_cfg(model.model_id, C.byref(cfg)))
        if precision is not None:  # default: f32 for L/NL, f64 for NL6 (mpcb_mppi_default_cfg)
            cfg.precision = A.PRECISIONS[precision]
        precision = {A.F32: "f32", A.F64: "f64", A.F64_FAST: "f64fast"}[cfg.precision]
        cfg.horizon, cfg.samples, cfg.state_dim = int(N), int(K), int(S)
        cfg.controllers, cfg.device, cfg.rank, cfg.world_size = int(controllers), int(device), int(rank), int(world_size)
        cfg.lambda_, cfg.std_dev = float(lam), float(std_dev)
        cfg.limit_lo, cfg.limit_hi = float(limit[0]), float(limit[1])
        if seed is not None:
            cfg.seed = int(seed)
        cfg.keep_costs = int(bool(keep_costs))
        if dt is not None:
            cfg.model.dt = float(dt)
        for k, v in ({} if isinstance(model, UserModel) else (params or {})).items():
            if k == "cost":
                for i, c in enumerate(v):
                    cfg.model.cost[i] = float(c)
            else:
                setattr(cfg.model, k, float(v))
        self.cfg = cfg
        self.N, self.K, self.S, self.C = int(N), int(K), int(S), int(controllers)
        self.precision = precision
        self._h = A._H()
        if isinstance(model, UserModel):
            pa = (C.c_double * max(1, len(model.params)))(*model.params)
            st = L.mpcb_mppi_create_user(C.byref(self._h), C.byref(cfg), model.source.encode(), pa, len(model.params))
            if st == A.RTC_ERROR:
                raise A.MpcB200Error(st, "user model did not compile: " + L.mpcb_last_error_string().decode() + "\n"
                                      + (L.mpcb_rtc_log() or b"").decode())
            A.check(st)
        else:
            A.check(L.mpcb_mppi_create(C.byref(self._h), C.byref(cfg)))
        self.K_local = L.mpcb_mppi_local_samples(self._h)
        self.info = None

    @classmethod
    def new(cls, dynamics: DeviceModel, cost: DeviceModel, lam: float, std_dev: float, limit, *, N: int, K: int,
            S: int = 4, **kw) -> "Mppi":
        """Argument order of Mppi::new (src/mppi.rs:16-22); N, K, S are the const generics."""
        if dynamics.model_id != cost.model_id:
            raise ValueError("dynamics and cost must name the same built-in device model")
        return cls(N, K, S, model=dynamics, lam=lam, std_dev=std_dev, limit=limit, **kw)

    # -- lifetime --
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value is not None:
            A.lib().mpcb_mppi_destroy(self._h)
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

    # -- helpers --
    def _inputs(self, x, u_n):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(self.C, self.S)
        u_n = np.ascontiguousarray(u_n, dtype=np.float64).reshape(self.C, self.N)
        return x, u_n

    def _info_dicts(self, infos):
        return [dict(status=i.status, argmax=i.argmax, max=i.max, sum=i.sum, n_finite=i.n_finite) for i in infos]

    def _finish(self, st, u_out, infos, squeeze):
        self.info = self._info_dicts(infos)
        if st in (A.NO_FINITE_COST, A.SUM_ZERO, A.U_INVALID):
            raise MppiError(st)
        A.check(st)
        return u_out[0] if squeeze else u_out

    def _fast_path(self):
        """Persistent host buffers + pre-bound pointers for compute(): the per-call Python cost is two small
        copies and one ctypes call."""
        fp = getattr(self, "_fp", None)
        if fp is None:
            x = np.zeros((self.C, self.S))
            u = np.zeros((self.C, self.N))
            out = np.zeros((self.C, self.N))
            infos = (A.MppiInfo * self.C)()
            fp = self._fp = (x, u, out, infos, _dp(x), _dp(u), _dp(out), A.lib().mpcb_mppi_compute)
        return fp

    # -- Mppi::compute (src/mppi.rs:33) --
    def compute(self, x, u_n):
        """Generate mode: noise drawn in-register (Philox).  Returns the new control sequence [N] (or [C][N])."""
        xb, ub, out, infos, px, pu, po, fn = self._fast_path()
        squeeze = self.C == 1 and np.ndim(u_n) == 1
        xb.reshape(-1)[:] = np.ravel(x)
        ub.reshape(-1)[:] = np.ravel(u_n)
        st = fn(self._h, px, pu, po, infos)
        self._last_infos = infos
        if st != A.OK:
            return self._finish(st, out.copy(), infos, squeeze)
        self.info = None  # filled lazily by last_call_info()
        return out[0].copy() if squeeze else out.copy()

    def last_call_info(self):
        """Diagnostics (status, argmax, max, sum, n_finite) of the last compute()."""
        if self.info is None and getattr(self, "_last_infos", None) is not None:
            self.info = self._info_dicts(self._last_infos)
        return self.info

    def compute_replay(self, x, u_n, eps):
        """Replay mode: eps[(C,)K,N] ~ N(0, std_dev^2) supplied by the caller (float32 or float64)."""
        squeeze = self.C == 1 and np.ndim(u_n) == 1
        x, u_n = self._inputs(x, u_n)
        eps = np.ascontiguousarray(eps)
        if eps.dtype not in (np.float32, np.float64):
            eps = eps.astype(np.float64)
        assert eps.size == self.C * self.K * self.N, "eps must hold all K global samples"
        u_out = np.empty((self.C, self.N))
        infos = (A.MppiInfo * self.C)()
        st = A.lib().mpcb_mppi_compute_replay(self._h, _dp(x), _dp(u_n), eps.ctypes.data_as(C.c_void_p),
                                              A.DT_F64 if eps.dtype == np.float64 else A.DT_F32, 0, _dp(u_out), infos)
        return self._finish(st, u_out, infos, squeeze)

    def compute_dump(self, x, u_n):
        """Generate mode that also returns the noise it drew: (u_out, eps[(C,)K_local,N])."""
        squeeze = self.C == 1 and np.ndim(u_n) == 1
        x, u_n = self._inputs(x, u_n)
        u_out = np.empty((self.C, self.N))
        infos = (A.MppiInfo * self.C)()
        eps = np.empty((self.C, self.K_local, self.N), dtype=np.float32 if self.precision == "f32" else np.float64)
        st = A.lib().mpcb_mppi_compute_dump(self._h, _dp(x), _dp(u_n), eps.ctypes.data_as(C.c_void_p), _dp(u_out), infos)
        u = self._finish(st, u_out, infos, squeeze)
        return u, (eps[0] if squeeze else eps)

    def costs(self):
        """c_k = -cost - control_term of the last compute (src/mppi.rs:61); needs keep_costs=True."""
        c = np.empty((self.C, self.K_local))
        A.check(A.lib().mpcb_mppi_get_costs(self._h, _dp(c)))
        return c[0] if self.C == 1 else c

    # -- device-resident / multi-GPU plumbing --
    def compute_device(self, d_x: int, d_u_in: int, d_u_out: int, d_eps: int = 0, eps_dtype: int = A.DT_F32):
        A.check(A.lib().mpcb_mppi_compute_device(self._h, d_x, d_u_in, d_eps or None, eps_dtype, d_u_out))

    def first_control_device(self, d_u_out: int, d_u0: int):
        """Asynchronous: d_u0[c] = d_u_out[c][0] (the control every controller applies next)."""
        A.check(A.lib().mpcb_mppi_first_control_device(self._h, d_u_out, d_u0))

    def sync(self):
        A.check(A.lib().mpcb_mppi_sync(self._h))

    def last_info(self):
        infos = (A.MppiInfo * self.C)()
        A.check(A.lib().mpcb_mppi_last_info(self._h, infos))
        return [dict(status=i.status, argmax=i.argmax, max=i.max, sum=i.sum, n_finite=i.n_finite) for i in infos]

    @property
    def stream(self) -> int:
        return A.lib().mpcb_mppi_stream(self._h) or 0

    @property
    def launches(self) -> int:
        return A.lib().mpcb_mppi_launches(self._h)

    @property
    def partial_len(self) -> int:
        return A.lib().mpcb_mppi_partial_len(self._h)

    def compute_partial(self, x, u_n, d_partial: int, d_eps: int = 0, eps_dtype: int = A.DT_F32):
        x, u_n = self._inputs(x, u_n)
        A.check(A.lib().mpcb_mppi_compute_partial(self._h, _dp(x), _dp(u_n), d_eps or None, eps_dtype, d_partial))

    def combine(self, d_partials: int, n_ranks: int):
        u_out = np.empty((self.C, self.N))
        infos = (A.MppiInfo * self.C)()
        st = A.lib().mpcb_mppi_combine(self._h, d_partials, n_ranks, _dp(u_out), infos)
        return self._finish(st, u_out, infos, self.C == 1)

    def attach_comm(self, unique_id: bytes):
        assert len(unique_id) == 128
        A.check(A.lib().mpcb_mppi_attach_comm(self._h, unique_id))


PEER_HANDLE_BYTES = 128


def _peer_handle(self) -> bytes:
    """128-byte handle of this rank's exchange mailbox (mpcb_mppi_peer_handle)."""
    buf = C.create_string_buffer(PEER_HANDLE_BYTES)
    A.check(A.lib().mpcb_mppi_peer_handle(self._h, buf))
    return buf.raw


def _attach_peers(self, handles):
    """handles: the peer_handle() of every rank, in rank order -> fused peer-memory exchange inside the kernel."""
    blob = b"".join(handles)
    assert len(blob) == PEER_HANDLE_BYTES * self.cfg.world_size
    A.check(A.lib().mpcb_mppi_attach_peers(self._h, blob))


Mppi.peer_handle = _peer_handle
Mppi.attach_peers = _attach_peers


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    A.check(A.lib().mpcb_comm_unique_id(buf))
    return buf.raw
