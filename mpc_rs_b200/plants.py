"""Host-side plants and sensors of the reference's example binaries (numpy, vectorised over a leading batch axis).

In the reference these live in the example files next to `main` — the simulated robot the controller/filter is
run against, not part of the library (`src/`).  They are restated here so the example drivers in `examples/` are
self-contained; the hot paths (MPPI rollouts, UKF) never call them.  Formula order follows the cited lines.
"""
from __future__ import annotations

import numpy as np

from . import _abi as A


def _params(model_id: int, **over):
    import ctypes as C
    p = A.ModelParams()
    A.check(A.lib().mpcb_model_defaults(model_id, C.byref(p)))
    for k, v in over.items():
        setattr(p, k, float(v))
    return p


class PlantL:
    """Linear cart-pendulum, semi-implicit Euler — examples/mppi4.rs:73-89 (DT = T/N = 0.1)."""

    def __init__(self, dt: float = 0.1):
        p = _params(A.MODEL_L)
        M1, R_W, M2, L, J1, J2, G, KT = p.m1, p.r_w, p.m2, p.l, p.j1, p.j2, p.g, p.kt
        D = (M1 + M2 + J1 / (R_W * R_W)) * (M2 * L * L + J2) - M2 * M2 * L * L
        self.a1 = (M1 + M2 + J1 / (R_W * R_W)) / D * M2 * G * L
        self.b1 = M2 * L / D / R_W * KT
        self.a2 = -M2 * M2 * G * L * L / D
        self.b2 = (M2 * L * L + J2) / D / R_W * KT
        self.dt = dt

    def step(self, x, u):
        r = np.array(x, dtype=np.float64, copy=True)
        r[..., 3] += (self.a1 * r[..., 2] - self.b1 * u) * self.dt
        r[..., 2] += r[..., 3] * self.dt
        r[..., 1] += (self.a2 * r[..., 2] + self.b2 * u) * self.dt
        r[..., 0] += r[..., 1] * self.dt
        return r


class PlantNL:
    """Nonlinear pendulum, explicit Euler on the old state — examples/mppi4-non-liner.rs:73-94."""

    def __init__(self, dt: float = 0.1):
        self.p = _params(A.MODEL_NL)
        self.dt = dt

    def step(self, x, u):
        p = self.p
        M1, R_W, M2, L, J1, J2, G, KT = p.m1, p.r_w, p.m2, p.l, p.j1, p.j2, p.g, p.kt
        x = np.asarray(x, dtype=np.float64)
        s, c = np.sin(x[..., 2]), np.cos(x[..., 2])
        D = (M1 + M2 + J1 / (R_W * R_W)) * (M2 * L * L + J2)
        d = D - M2 * M2 * L * L * c * c
        term1 = (M1 + M2 + J1 / (R_W * R_W)) * M2 * G * L * s
        q = KT * u / R_W + M2 * L * (x[..., 3] * x[..., 3]) * s
        term2 = q * M2 * L * c
        r = np.empty_like(x)
        r[..., 3] = x[..., 3] + (term1 - term2) / d * self.dt
        r[..., 2] = x[..., 2] + x[..., 3] * self.dt
        term3 = (J2 + M2 * L * L) * q
        term4 = M2 * G * L * L * s * c
        r[..., 1] = x[..., 1] + (term3 + term4) / d * self.dt
        r[..., 0] = x[..., 0] + x[..., 1] * self.dt
        return r


class PlantPenLin:
    """Truth model + sensor of examples/ukf-pen.rs:76-118 (the same linear model with J2 = 0.1, DT = 0.01)."""

    R_DIAG = (0.5, 0.5)  # :23-26; the sensor uses these as standard deviations (:114)

    def __init__(self):
        p = _params(A.MODEL_PEN_LIN)
        M1, R_W, M2, L, J1, J2, G, KT = p.m1, p.r_w, p.m2, p.l, p.j1, p.j2, p.g, p.kt
        D = (M1 + M2 + J1 / (R_W * R_W)) * (M2 * L * L + J2) - M2 * M2 * L * L
        self.a1 = (M1 + M2 + J1 / (R_W * R_W)) / D * M2 * G * L
        self.b1 = M2 * L / D / R_W * KT
        self.a2 = -M2 * M2 * G * L * L / D
        self.b2 = (M2 * L * L + J2) / D / R_W * KT
        self.dt = p.dt

    def fx(self, x, u):
        r = np.array(x, dtype=np.float64, copy=True)
        r[..., 3] += (self.a1 * r[..., 2] - self.b1 * u) * self.dt
        r[..., 2] += r[..., 3] * self.dt
        r[..., 1] += (self.a2 * r[..., 2] + self.b2 * u) * self.dt
        r[..., 0] += r[..., 1] * self.dt
        return r

    def sensor(self, x, rng):
        z = np.stack([x[..., 1], x[..., 3]], axis=-1)
        return z + np.asarray(self.R_DIAG) * rng.standard_normal(z.shape)


class PlantPenNL:
    """Truth model + sensor of examples/ukf-pen2.rs:31-66 (nonlinear pendulum, DT = 0.01; o = 3: rpm, rpm, deg/s)."""

    NOISE = (100.0, 100.0, 0.5)  # :58-64

    def __init__(self):
        self.p = _params(A.MODEL_PEN_NL)
        self.dt = self.p.dt

    def fx(self, x, u):
        p = self.p
        M1, R_W, M2, L, J1, J2, G, KT = p.m1, p.r_w, p.m2, p.l, p.j1, p.j2, p.g, p.kt
        x = np.asarray(x, dtype=np.float64)
        th, om = x[..., 2], x[..., 3]
        D = (M1 + M2 + J1 / (R_W * R_W)) * (M2 * L * L + J2)
        d = D - M2 * M2 * L * L * np.cos(th) * np.cos(th)
        drive = KT * u / R_W + M2 * L * om ** 2 * np.sin(th)
        r = np.array(x, copy=True)
        r[..., 3] += ((M1 + M2 + J1 / (R_W * R_W)) * M2 * G * L * np.sin(th) - drive * M2 * L * np.cos(th)) / d * self.dt
        r[..., 2] += om * self.dt
        r[..., 1] += ((J2 + M2 * L * L) * drive + M2 * G * L * L * np.sin(th) * np.cos(th)) / d * self.dt
        r[..., 0] += x[..., 1] * self.dt
        return r

    def hx(self, x):
        rpm = 60.0 / (2.0 * np.pi * self.p.r_w) * x[..., 1]
        return np.stack([rpm, rpm, np.degrees(x[..., 3])], axis=-1)

    def sensor(self, x, rng):
        z = self.hx(np.asarray(x, dtype=np.float64))
        return z + np.asarray(self.NOISE) * rng.standard_normal(z.shape)


class PlantPen6:
    """Truth model + sensor of examples/ukf-pen3.rs:35-77: state [x, x', x'', th, th', th''], accelerations recomputed
    every step (the `x[2].cos()` in the denominator is the reference's, :38); o = 5."""

    NOISE = (100.0, 100.0, 0.5, 100.0, 100.0)  # :68-74

    def __init__(self):
        self.p = _params(A.MODEL_PEN6)
        self.dt = self.p.dt

    def fx(self, x, u):
        p = self.p
        M1, R_W, M2, L, J1, J2, G, KT = p.m1, p.r_w, p.m2, p.l, p.j1, p.j2, p.g, p.kt
        x = np.asarray(x, dtype=np.float64)
        th, om = x[..., 3], x[..., 4]
        D = (M1 + M2 + J1 / (R_W * R_W)) * (M2 * L * L + J2)
        d = D - (M2 * L * np.cos(x[..., 2])) ** 2
        drive = KT * u / R_W + M2 * L * om ** 2 * np.sin(th)
        r = np.array(x, copy=True)
        r[..., 0] += x[..., 1] * self.dt
        r[..., 1] += x[..., 2] * self.dt
        r[..., 2] = ((J2 + M2 * L * L) * drive + M2 * G * L * L * np.sin(th) * np.cos(th)) / d
        r[..., 3] += x[..., 4] * self.dt
        r[..., 4] += x[..., 5] * self.dt
        r[..., 5] = ((M1 + M2 + J1 / (R_W * R_W)) * M2 * G * L * np.sin(th) - drive * M2 * L * np.cos(th)) / d
        return r

    def hx(self, x):
        p = self.p
        M2, L, G = p.m2, p.l, p.g
        th = x[..., 3]
        v = M2 * G * np.cos(th) + M2 * x[..., 2] * np.sin(th) - M2 * L * x[..., 4] ** 2
        h = -M2 * G * np.sin(th) + M2 * x[..., 2] * np.cos(th) + M2 * L * x[..., 5]
        rpm = 60.0 / (2.0 * np.pi * p.r_w) * x[..., 1]
        return np.stack([rpm, rpm, np.degrees(th), v / G, h / G], axis=-1)

    def sensor(self, x, rng):
        z = self.hx(np.asarray(x, dtype=np.float64))
        return z + np.asarray(self.NOISE) * rng.standard_normal(z.shape)


class PlantNL6:
    """The robot of examples/mppi4-non-liner-ukf.rs: ddot :126-139, dynamics_short :149-159, hx :169-179,
    sensor :180-190 (R entries used as standard deviations), 2 N push for 1.0 < t < 1.5 (:236-241)."""

    R_DIAG = (200.0, 200.0, 10.0, 0.05, 0.05)  # :28

    def __init__(self):
        self.p = _params(A.MODEL_NL6)

    def ddot(self, x4, u, f):
        p = self.p
        M1, R_W, M2, L, J1, J2, G, KT = p.m1, p.r_w, p.m2, p.l, p.j1, p.j2, p.g, p.kt
        D1 = (2.0 * M1 + M2 + 2.0 * J1 / (R_W * R_W)) * (M2 * L * L + J2)
        th, thd = x4[..., 2], x4[..., 3]
        d = D1 - (M2 * L * np.cos(th)) ** 2
        term1 = (M2 * L * L + J2) * M2 * L / d * thd ** 2 * np.sin(th)
        term2 = -(M2 * L) ** 2 * G / d * np.sin(th) * np.cos(th)
        term3 = 2.0 * (M2 * L * L + J2) / (d * R_W) * KT * u
        term4 = (M2 * L * L + J2) / d * f * np.cos(thd)  # x[3].cos() as written (:131)
        ddx = term1 + term2 + term3 + term4
        t1 = -(M2 * L) ** 2 / d * thd ** 2 * np.sin(th) * np.cos(th)
        t2 = (M2 * G * np.sin(th) - 2.0 * f) * L * (2.0 * M1 + M2 + 2.0 * J1 / (R_W * R_W)) / d
        t3 = -2.0 * M2 * L / (d * R_W) * KT * u * np.cos(th)
        t4 = -M2 * L * f * np.cos(thd) ** 2 / d
        return ddx, t1 + t2 + t3 + t4

    def dynamics_short(self, x6, u, dt, f=0.0):
        x6 = np.asarray(x6, dtype=np.float64)
        ddx, ddth = self.ddot(np.stack([x6[..., 0], x6[..., 1], x6[..., 3], x6[..., 4]], axis=-1), u, f)
        r = x6.copy()
        r[..., 5] = ddth
        r[..., 4] += r[..., 5] * dt
        r[..., 3] += r[..., 4] * dt
        r[..., 2] = ddx
        r[..., 1] += r[..., 2] * dt
        r[..., 0] += r[..., 1] * dt
        return r

    def hx(self, s):
        p = self.p
        G, L, R_W = p.g, p.l, p.r_w
        ax = G * np.sin(s[..., 3]) + s[..., 2] * np.cos(s[..., 3]) + L * s[..., 5]
        az = G * np.cos(s[..., 3]) - s[..., 2] * np.sin(s[..., 3]) + L * s[..., 4] ** 2
        return np.stack([36.0 * 60.0 / (2.0 * np.pi * R_W) * s[..., 1], 36.0 * -60.0 / (2.0 * np.pi * R_W) * s[..., 1],
                         np.degrees(s[..., 4]), az / G, ax / G], axis=-1)

    def sensor(self, x6, rng):
        z = self.hx(x6)
        return z + np.asarray(self.R_DIAG) * rng.standard_normal(z.shape)

    @staticmethod
    def push(t: float) -> float:
        return 2.0 if 1.0 < t < 1.5 else 0.0
