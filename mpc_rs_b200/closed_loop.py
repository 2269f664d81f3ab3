"""Batched closed loop of examples/mppi4-non-liner-ukf.rs (BASELINE config #4): C independent robots, each with
its own MPPI controller (model NL6, `controllers = C` in one launch) and its own UKF (model NL6_UKF, one filter
per robot), coupled on the device.

The reference runs four free-running threads against the wall clock (plant :224-246, UKF :248-292, MPPI :51-103,
log :391-432).  A batch needs a deterministic schedule, so one `tick()` of fixed length `tick_dt` (SURVEY.md 8d:
0.01 s, the reference's ~9-10 ms sensor period :267-268) does, in the order the data flows in the example:

    plant    x   <- dynamics_short(x, u_n[0], tick_dt, push(t))            host, numpy      (:236-244)
    sensor   z   <- hx(x) + R * N(0,1)                                     host, numpy      (:180-190)
    UKF      set_q(gen_q(dt)); predict(u_n[0], fx); update(z, hx)          GPU, one launch  (:272-283)
    MPPI     x_est -> [x0, x1, x3, x4]; u_n <- compute(x_est, u_n)         GPU, one launch  (:55-87)

The estimate goes from the UKF's device state to the MPPI input, and the new u_n[0] back to the UKF's control
input, without leaving the device (mpcb_ukf_gather_state_device / mpcb_mppi_first_control_device); only z (5 values
per robot) goes up and u_n[0] (1 value) comes down per tick, because the plant stands in for the real robot.
MPPI failures zero that robot's control sequence like the example's `Err(e) => zeros` (:81-86).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi as A
from . import models
from .mppi import Mppi
from .plants import PlantNL6
from .ukf import BatchedUkf, default_noise


def _dev_alloc(dev: int, nbytes: int) -> int:
    p = C.c_void_p()
    A.check(A.lib().mpcb_device_alloc(dev, nbytes, C.byref(p)))
    return p.value


def _upload(dev: int, dptr: int, arr: np.ndarray):
    arr = np.ascontiguousarray(arr)
    A.check(A.lib().mpcb_device_upload(dev, dptr, arr.ctypes.data_as(C.c_void_p), arr.nbytes))


def _download(dev: int, dptr: int, arr: np.ndarray):
    A.check(A.lib().mpcb_device_download(dev, arr.ctypes.data_as(C.c_void_p), dptr, arr.nbytes))


class ClosedLoopBatch:
    # constants of examples/mppi4-non-liner-ukf.rs:8-28
    T, N = 1.2, 8
    LAMBDA, R_U, LIMIT = 1.4, 4.0, (-10.0, 10.0)

    def __init__(self, controllers: int, samples: int = 8192, *, tick_dt: float = 0.01, use_estimate: bool = True,
                 precision: str = None, exact_ukf: bool = False, seed: int = 20240004, device: int = 0, x0=None):
        self.C, self.K, self.H = int(controllers), int(samples), self.N
        self.DT = self.T / self.N
        self.tick_dt, self.use_estimate, self.dev = float(tick_dt), bool(use_estimate), int(device)
        self.plant = PlantNL6()
        self.rng = np.random.Generator(np.random.PCG64(seed))
        self.mppi = Mppi(self.H, self.K, model=models.NL6, lam=self.LAMBDA, std_dev=self.R_U, limit=self.LIMIT,
                         precision=precision, controllers=self.C, seed=seed, device=device, dt=self.DT)
        self.ukf = BatchedUkf(models.NL6_UKF, self.C, device=device, exact=exact_ukf)
        Q, R, P0 = default_noise(models.NL6_UKF, self.tick_dt)  # gen_q(dt) :192-221, R :28, P0 = 10 I :163
        self.x = np.zeros((self.C, 6)) if x0 is None else np.array(x0, dtype=np.float64).reshape(self.C, 6)
        self.ukf.init(np.zeros(6), P0, Q, R)
        self.ukf.set_state(self.x, None)  # init_ukf(&init_x), :40,161-167
        self.u0 = np.zeros(self.C)
        self.t = 0.0
        self.ticks = 0
        n = self.C
        self.d_x4 = _dev_alloc(device, 8 * 4 * n)
        self.d_u = [_dev_alloc(device, 8 * self.H * n), _dev_alloc(device, 8 * self.H * n)]
        self.d_u0 = _dev_alloc(device, 8 * n)
        self.d_z = _dev_alloc(device, 8 * 5 * n)
        self.d_eps = 0
        _upload(device, self.d_u[0], np.zeros((n, self.H)))
        _upload(device, self.d_u0, self.u0)
        self.cur = 0

    def close(self):
        for p in (self.d_x4, *self.d_u, self.d_u0, self.d_z, self.d_eps):
            if p:
                A.lib().mpcb_device_free(self.dev, p)
        self.d_x4 = self.d_u0 = self.d_z = self.d_eps = 0
        self.d_u = [0, 0]
        self.mppi.close()
        self.ukf.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def tick(self, eps=None, z=None):
        """One tick.  eps[C][K][H] replays given MPPI noise (verification); z[C][5] replaces the simulated sensor."""
        dev = self.dev
        # plant + sensor (host)
        self.x = self.plant.dynamics_short(self.x, self.u0, self.tick_dt, self.plant.push(self.t))
        self.z = self.plant.sensor(self.x, self.rng) if z is None else np.asarray(z, dtype=np.float64).reshape(self.C, 5)
        # UKF: fused predict(u_n[0]) + update(z) with the per-robot control that is being applied
        _upload(dev, self.d_z, self.z.T)  # SoA [5][C]
        self.ukf.run_device(1, self.d_z, d_u=self.d_u0, dt=self.tick_dt)
        if self.use_estimate:
            self.ukf.gather_state_device((0, 1, 3, 4), self.d_x4)  # :78
            self.ukf.sync()
        else:  # DEBUG_UKF = true: the controller sees the true state (:55-57)
            self.ukf.sync()
            _upload(dev, self.d_x4, self.x[:, (0, 1, 3, 4)])
        # MPPI: all C controllers in one launch, previous sequence in, new sequence out
        nxt = self.cur ^ 1
        d_eps, eps_dt = 0, A.DT_F32
        if eps is not None:
            eps = np.ascontiguousarray(eps)
            assert eps.shape == (self.C, self.K, self.H) and eps.dtype in (np.float32, np.float64)
            if not self.d_eps:
                self.d_eps = _dev_alloc(dev, 8 * eps.size)
            _upload(dev, self.d_eps, eps)
            d_eps, eps_dt = self.d_eps, (A.DT_F64 if eps.dtype == np.float64 else A.DT_F32)
        self.mppi.compute_device(self.d_x4, self.d_u[self.cur], self.d_u[nxt], d_eps=d_eps, eps_dtype=eps_dt)
        self.mppi.first_control_device(self.d_u[nxt], self.d_u0)
        self.mppi.sync()
        self.cur = nxt
        _download(dev, self.d_u0, self.u0)
        self.t += self.tick_dt
        self.ticks += 1
        return self.u0

    # -- inspection --
    def controls(self):
        u = np.empty((self.C, self.H))
        _download(self.dev, self.d_u[self.cur], u)
        return u

    def estimate(self):
        return self.ukf.get_state()

    def mppi_status(self):
        return np.array([i["status"] for i in self.mppi.last_info()])

    def upright(self):
        """Robots whose true pitch is still inside the reference's abort bound |theta| <= pi/2 (:61)."""
        return np.abs(self.x[:, 3]) <= np.pi / 2
