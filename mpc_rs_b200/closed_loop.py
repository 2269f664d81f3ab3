"""Batched closed loop of examples/mppi4-non-liner-ukf.rs (BASELINE config #4): C independent robots, each with
its own MPPI controller (model NL6, `controllers = C` in one launch) and its own UKF (model NL6_UKF, one filter
per robot), coupled on the device — host-side mirror of `mpcb_closed_loop_*` (include/mpc_b200.h, csrc/closed_loop.cu).

The reference runs four free-running threads against the wall clock (plant :224-246, UKF :248-292, MPPI :51-103,
log :391-432).  A batch needs a deterministic schedule, so one tick of fixed length `tick_dt` (SURVEY.md 8d:
0.01 s, the reference's ~9-10 ms sensor period :267-268) does, in the order the data flows in the example:

    plant    x   <- dynamics_short(x, u_n[0], tick_dt, push(t))            GPU, one launch  (:236-244)
    sensor   z   <- hx(x) + R * N(0,1)                                     the same launch  (:180-190)
    UKF      set_q(gen_q(dt)); predict(u_n[0], fx); update(z, hx)          GPU, one launch  (:272-283)
    MPPI     x_est -> [x0, x1, x3, x4]; u_n <- compute(x_est, u_n)         GPU, one launch  (:55-87)

Everything stays on the device (round 1 ran the plant and the sensor in numpy on the host: 35 % of a tick); `tick(n)`
only enqueues 5 n launches.  MPPI failures zero that robot's control sequence like the example's `Err(e) => zeros`
(:81-86).  Robots shard over GPUs without any exchange: `controller_offset` = this handle's first global robot.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi as A
from .plants import PlantNL6


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class ClosedLoopBatch:
    # constants of examples/mppi4-non-liner-ukf.rs:8-28
    T, N = 1.2, 8
    LAMBDA, R_U, LIMIT = 1.4, 4.0, (-10.0, 10.0)

    def __init__(self, controllers: int, samples: int = 8192, *, tick_dt: float = 0.01, use_estimate: bool = True,
                 precision: str = None, exact_ukf: bool = False, seed: int = 20240004, device: int = 0, x0=None,
                 controller_offset: int = 0):
        L = A.lib()
        self.C, self.K, self.H = int(controllers), int(samples), self.N
        self.DT = self.T / self.N
        self.tick_dt, self.use_estimate, self.dev = float(tick_dt), bool(use_estimate), int(device)
        self.plant = PlantNL6()  # host restatement of the robot (logging / prediction lines of the example only)
        cfg = A.ClosedLoopCfg()
        A.check(L.mpcb_closed_loop_default_cfg(C.byref(cfg)))
        cfg.controllers, cfg.samples, cfg.controller_offset = self.C, self.K, int(controller_offset)
        cfg.tick_dt, cfg.seed, cfg.use_estimate = self.tick_dt, int(seed), int(self.use_estimate)
        cfg.precision = A.PRECISIONS.get(precision, -1)
        cfg.exact_ukf, cfg.device = int(bool(exact_ukf)), self.dev
        self._h = A._H()
        A.check(L.mpcb_closed_loop_create(C.byref(self._h), C.byref(cfg)))
        if x0 is not None:
            self.set_state(np.array(x0, dtype=np.float64).reshape(self.C, 6))
        self.z = np.zeros((self.C, 5))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value is not None:
            A.lib().mpcb_closed_loop_destroy(self._h)
            self._h = A._H()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- state --
    def set_state(self, x6, P=None):
        """Truth and estimate of every robot (init_ukf(&init_x), :40,161-167); P[C][6][6] keeps P0 when None."""
        x6 = np.ascontiguousarray(x6, dtype=np.float64).reshape(self.C, 6)
        P = None if P is None else np.ascontiguousarray(P, dtype=np.float64).reshape(self.C, 6, 6)
        A.check(A.lib().mpcb_closed_loop_set_state(self._h, _dp(x6), _dp(P)))

    def set_truth(self, x6):
        x6 = np.ascontiguousarray(x6, dtype=np.float64).reshape(self.C, 6)
        A.check(A.lib().mpcb_closed_loop_set_truth(self._h, _dp(x6)))

    def set_estimate(self, x_est, P):
        x_est = np.ascontiguousarray(x_est, dtype=np.float64).reshape(self.C, 6)
        P = np.ascontiguousarray(P, dtype=np.float64).reshape(self.C, 6, 6)
        A.check(A.lib().mpcb_ukf_set_state(A.lib().mpcb_closed_loop_ukf(self._h), _dp(x_est), _dp(P)))

    def set_controls(self, u_seq):
        u_seq = np.ascontiguousarray(u_seq, dtype=np.float64).reshape(self.C, self.H)
        A.check(A.lib().mpcb_closed_loop_set_controls(self._h, _dp(u_seq)))

    # -- ticks --
    def tick(self, n: int = 1, *, eps=None, z=None):
        """n ticks (asynchronous).  With eps[C][K][H] (replayed MPPI noise) and/or z[C][5] (sensor readings instead of the
        simulated sensor): one synchronous verification tick; returns the applied controls u_n[0] of every robot then."""
        if eps is None and z is None:
            A.check(A.lib().mpcb_closed_loop_tick(self._h, int(n)))
            return None
        assert n == 1
        zp, ep, edt = None, None, A.DT_F32
        if z is not None:
            z = np.ascontiguousarray(z, dtype=np.float64).reshape(self.C, 5)
            zp = _dp(z)
        if eps is not None:
            eps = np.ascontiguousarray(eps)
            assert eps.shape == (self.C, self.K, self.H) and eps.dtype in (np.float32, np.float64)
            ep, edt = eps.ctypes.data_as(C.c_void_p), (A.DT_F64 if eps.dtype == np.float64 else A.DT_F32)
        A.check(A.lib().mpcb_closed_loop_tick_replay(self._h, zp, ep, edt))
        return self.applied()

    def sync(self):
        A.check(A.lib().mpcb_closed_loop_sync(self._h))

    @property
    def ticks(self) -> int:
        return int(A.lib().mpcb_closed_loop_ticks(self._h))

    @property
    def t(self) -> float:
        return self.ticks * self.tick_dt

    @property
    def launches(self) -> int:
        return int(A.lib().mpcb_closed_loop_launches(self._h))

    # -- inspection (each call synchronises) --
    def _get(self, **want):
        bufs = {"x6": (self.C, 6), "x_est": (self.C, 6), "z": (self.C, 5), "u0": (self.C,), "u_seq": (self.C, self.H)}
        out = {k: (np.empty(shape) if want.get(k) else None) for k, shape in bufs.items()}
        st = (C.c_int32 * self.C)() if want.get("status") else None
        A.check(A.lib().mpcb_closed_loop_get(self._h, _dp(out["x6"]), _dp(out["x_est"]), _dp(out["z"]), _dp(out["u0"]), _dp(out["u_seq"]), st))
        if st is not None:
            out["status"] = np.array(st[:], dtype=np.int32)
        return out

    @property
    def x(self):
        """True state of every robot, [C][6]."""
        return self._get(x6=True)["x6"]

    def readings(self):
        return self._get(z=True)["z"]

    def applied(self):
        return self._get(u0=True)["u0"]

    def controls(self):
        return self._get(u_seq=True)["u_seq"]

    def estimate(self):
        """(x_est[C][6], P[C][6][6]) of the filters."""
        L = A.lib()
        x, P = np.empty((self.C, 6)), np.empty((self.C, 6, 6))
        self.sync()
        A.check(L.mpcb_ukf_get_state(L.mpcb_closed_loop_ukf(self._h), _dp(x), _dp(P)))
        return x, P

    def estimate_range(self, first: int, count: int):
        L = A.lib()
        x, P = np.empty((count, 6)), np.empty((count, 6, 6))
        self.sync()
        A.check(L.mpcb_ukf_get_state_range(L.mpcb_closed_loop_ukf(self._h), int(first), int(count), _dp(x), _dp(P)))
        return x, P

    def mppi_status(self):
        return self._get(status=True)["status"]

    def upright(self):
        """Robots whose true pitch is still inside the reference's abort bound |theta| <= pi/2 (:61)."""
        return np.abs(self.x[:, 3]) <= np.pi / 2
