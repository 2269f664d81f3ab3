"""Independent numpy re-derivation of the reference algorithm, written from the Rust sources separately from
oracle/mpc_oracle.c.  TEST INFRASTRUCTURE ONLY: it cross-checks the C oracle (the reference itself cannot be
built here — no Rust toolchain — and ships no golden vectors).  Vectorised over samples, library
linear algebra (np.linalg) instead of hand-written Cholesky/Jacobi/LU, so it shares no code path with the oracle.
"""
import numpy as np

# ---- constants of the examples -------------------------------------------------------------------------
M1, R_W, L_, G, KT = 150e-3, 50e-3, 0.2474, 9.81, 0.15
M2 = 2.3 - 2.0 * M1 + 2.0
J1 = M1 * R_W * R_W


def cost_clamped(x):  # examples/mppi4.rs:20-27
    xc = np.clip(x[..., 0], -2.0, 2.0)
    return (2.0 * xc ** 2 + 3.0 * np.clip(x[..., 1] + 2.0 * xc, -5.0, 5.0) ** 2
            + 5.0 * (x[..., 2] + 0.35 * np.clip(x[..., 0], -0.75, 0.75)) ** 2 + 1.2 * x[..., 3] ** 2)


def dyn_L(x, u, dt, J2=0.2):  # examples/mppi4.rs:81-89
    D = (M1 + M2 + J1 / (R_W * R_W)) * (M2 * L_ * L_ + J2) - M2 * M2 * L_ * L_
    x = np.array(x, dtype=np.float64, copy=True)
    x[..., 3] += ((M1 + M2 + J1 / (R_W * R_W)) / D * M2 * G * L_ * x[..., 2] - M2 * L_ / D / R_W * KT * u) * dt
    x[..., 2] += x[..., 3] * dt
    x[..., 1] += (-M2 * M2 * G * L_ * L_ / D * x[..., 2] + (M2 * L_ * L_ + J2) / D / R_W * KT * u) * dt
    x[..., 0] += x[..., 1] * dt
    return x


def dyn_NL(x, u, dt, J2=0.2):  # examples/mppi4-non-liner.rs:81-94
    x = np.asarray(x, dtype=np.float64)
    r = x.copy()
    D = (M1 + M2 + J1 / (R_W * R_W)) * (M2 * L_ * L_ + J2)
    c, s = np.cos(x[..., 2]), np.sin(x[..., 2])
    d = D - M2 * M2 * L_ * L_ * c * c
    term1 = (M1 + M2 + J1 / (R_W * R_W)) * M2 * G * L_ * s
    q = KT * u / R_W + M2 * L_ * x[..., 3] ** 2 * s
    term2 = q * M2 * L_ * c
    r[..., 3] += (term1 - term2) / d * dt
    r[..., 2] += x[..., 3] * dt
    term3 = (J2 + M2 * L_ * L_) * q
    term4 = M2 * G * L_ * L_ * s * c
    r[..., 1] += (term3 + term4) / d * dt
    r[..., 0] += x[..., 1] * dt
    return r


# examples/mppi4-non-liner-ukf.rs:108-124
N6 = dict(M1=160e-3, R_W=50e-3, M2=2.4, L=0.4, J1=2.23e5 * 1e-9, J2=1.168e8 * 1e-9, G=9.81, KT=0.15)


def ddot6(x4, u, f):  # :126-139
    c = N6
    m1, rw, m2, l, j1, j2, g, kt = c["M1"], c["R_W"], c["M2"], c["L"], c["J1"], c["J2"], c["G"], c["KT"]
    D1 = (2.0 * m1 + m2 + 2.0 * j1 / (rw * rw)) * (m2 * l * l + j2)
    th, thd = x4[..., 2], x4[..., 3]
    d = D1 - (m2 * l * np.cos(th)) ** 2
    ddx = ((m2 * l * l + j2) * m2 * l / d * thd ** 2 * np.sin(th) - (m2 * l) ** 2 * g / d * np.sin(th) * np.cos(th)
           + 2.0 * (m2 * l * l + j2) / (d * rw) * kt * u + (m2 * l * l + j2) / d * f * np.cos(thd))
    ddth = (-(m2 * l) ** 2 / d * thd ** 2 * np.sin(th) * np.cos(th)
            + (m2 * g * np.sin(th) - 2.0 * f) * l * (2.0 * m1 + m2 + 2.0 * j1 / (rw * rw)) / d
            - 2.0 * m2 * l / (d * rw) * kt * u * np.cos(th) - m2 * l * f * np.cos(thd) ** 2 / d)
    return ddx, ddth


def dyn_NL6(x, u, dt):  # dynamics4 :140-148
    x = np.asarray(x, dtype=np.float64)
    ddx, ddth = ddot6(x, u, 0.0)
    r = x.copy()
    r[..., 3] += ddth * dt
    r[..., 2] += r[..., 3] * dt
    r[..., 1] += ddx * dt
    r[..., 0] += r[..., 1] * dt
    return r


def cost_quadratic(x):  # :33-35
    return 0.1 * x[..., 0] ** 2 + 0.1 * x[..., 1] ** 2 + 1.0 * x[..., 2] ** 2 + 0.5 * x[..., 3] ** 2


def dynamics_short(x6, u, dt, f):  # :149-159
    x6 = np.asarray(x6, dtype=np.float64)
    ddx, ddth = ddot6(np.stack([x6[..., 0], x6[..., 1], x6[..., 3], x6[..., 4]], -1), u, f)
    r = x6.copy()
    r[..., 5] = ddth
    r[..., 4] += r[..., 5] * dt
    r[..., 3] += r[..., 4] * dt
    r[..., 2] = ddx
    r[..., 1] += r[..., 2] * dt
    r[..., 0] += r[..., 1] * dt
    return r


MPPI_MODELS = {0: (dyn_L, cost_clamped), 1: (dyn_NL, cost_clamped), 2: (dyn_NL6, cost_quadratic)}


def mppi_compute(model_id, dt, lam, std_dev, lim, x, u_n, eps):
    """src/mppi.rs:33-92 on supplied noise eps[K][H]; returns (u_new, c[K]) or raises ValueError(message)."""
    dyn, cost = MPPI_MODELS[model_id]
    K, H = eps.shape
    v = np.clip(u_n[None, :] + eps, lim[0], lim[1])
    inv = std_dev ** -2
    xs = np.tile(np.asarray(x, dtype=np.float64), (K, 1))
    J = np.zeros(K)
    for t in range(H):
        xs = dyn(xs, v[:, t], dt)
        J = J + cost(xs)
    ctrl = np.zeros(K)
    for t in range(H):
        ctrl = ctrl + u_n[t] * inv * v[:, t]
    c = -J - ctrl
    fin = np.isfinite(c)
    if not fin.any():
        raise ValueError("Cannot calculate max")
    mx = c[fin].max()
    w = np.exp((c - mx) / lam)
    s = w.sum()
    if s == 0.0:
        raise ValueError("sum is zero")
    u_new = ((w / s)[:, None] * v).sum(0)
    if not np.isfinite(u_new[0]):
        raise ValueError("u is invalid")
    return u_new, c


# ---- UKF (src/ukf.rs) ------------------------------------------------------------------------------------
def ukf_weights(n):
    alpha, beta = 1e-3, 2.0
    kappa = 3.0 - n
    C = alpha * alpha * (n + kappa)
    lam = C - n
    wm = np.full(2 * n + 1, 1.0 / (2.0 * C))
    wc = wm.copy()
    wm[0] = lam / C
    wc[0] = lam / C + 1.0 - alpha ** 2 + beta
    return wm, wc, C


def sigma_points(x, P, sqrt_mode, order):
    n = len(x)
    _, _, C = ukf_weights(n)
    if sqrt_mode == "cholesky":
        Lm = np.linalg.cholesky(C * P)
    else:  # U sqrt(S) of the SVD of a symmetric PSD matrix
        w, V = np.linalg.eigh(C * (P + P.T) / 2)
        Lm = V * np.sqrt(np.abs(w))
    cols = [x]
    if order == "interleaved":
        for i in range(n):
            cols += [x + Lm[:, i], x - Lm[:, i]]
    else:
        cols += [x + Lm[:, i] for i in range(n)] + [x - Lm[:, i] for i in range(n)]
    return np.stack(cols, 1)


def unscented_transform(sig, wm, wc, cov):
    mean = sig @ wm
    y = sig - mean[:, None]
    return mean, (wc * y) @ y.T + cov


def ukf_predict(fx, x, P, Q, sqrt_mode="svd", order="library"):
    n = len(x)
    wm, wc, _ = ukf_weights(n)
    sig = sigma_points(x, P, sqrt_mode, order)
    sig = np.stack([fx(sig[:, i]) for i in range(sig.shape[1])], 1)
    x, P = unscented_transform(sig, wm, wc, Q)
    return x, P, sig


def ukf_update(hx, x, P, R, z, sig):
    n = len(x)
    wm, wc, _ = ukf_weights(n)
    zs = np.stack([hx(sig[:, i]) for i in range(sig.shape[1])], 1)
    zp, pz = unscented_transform(zs, wm, wc, R)
    pxz = (wc * (sig - x[:, None])) @ (zs - zp[:, None]).T
    k = pxz @ np.linalg.inv(pz)
    x = x + k @ (z - zp)
    P = P - k @ pz @ k.T
    return x, (P + P.T) / 2.0


def fx_pen_lin(x, u, dt=0.01):  # examples/ukf-pen.rs:76-83 (J2 = 0.1)
    return dyn_L(x, u, dt, J2=0.1)


def hx_pen_lin(x):
    return np.array([x[1], x[3]])


def fx_pen_nl(x, u, dt=0.01):  # examples/ukf-pen2.rs:31-44
    return dyn_NL(x, u, dt)


def hx_pen_nl(x):  # :47-53
    return np.array([60.0 / (2.0 * np.pi * R_W) * x[1], 60.0 / (2.0 * np.pi * R_W) * x[1], np.degrees(x[3])])


def fx_pen6(x, u, dt=0.01):  # examples/ukf-pen3.rs:35-50
    J2 = 0.2
    r = np.array(x, dtype=np.float64, copy=True)
    D = (M1 + M2 + J1 / (R_W * R_W)) * (M2 * L_ * L_ + J2)
    d = D - (M2 * L_ * np.cos(x[2])) ** 2
    r[0] += x[1] * dt
    r[1] += x[2] * dt
    q = KT * u / R_W + M2 * L_ * x[4] ** 2 * np.sin(x[3])
    r[2] = ((J2 + M2 * L_ * L_) * q + M2 * G * L_ * L_ * np.sin(x[3]) * np.cos(x[3])) / d
    r[3] += x[4] * dt
    r[4] += x[5] * dt
    r[5] = ((M1 + M2 + J1 / (R_W * R_W)) * M2 * G * L_ * np.sin(x[3]) - q * M2 * L_ * np.cos(x[3])) / d
    return r


def hx_pen6(x):  # :53-63
    v = M2 * G * np.cos(x[3]) + M2 * x[2] * np.sin(x[3]) - M2 * L_ * x[4] ** 2
    h = -M2 * G * np.sin(x[3]) + M2 * x[2] * np.cos(x[3]) + M2 * L_ * x[5]
    k = 60.0 / (2.0 * np.pi * R_W)
    return np.array([k * x[1], k * x[1], np.degrees(x[3]), v / G, h / G])


def hx_nl6(x):  # examples/mppi4-non-liner-ukf.rs:169-179
    g, l, rw = N6["G"], N6["L"], N6["R_W"]
    ax = g * np.sin(x[3]) + x[2] * np.cos(x[3]) + l * x[5]
    az = g * np.cos(x[3]) - x[2] * np.sin(x[3]) + l * x[4] ** 2
    return np.array([36.0 * 60.0 / (2.0 * np.pi * rw) * x[1], 36.0 * -60.0 / (2.0 * np.pi * rw) * x[1],
                     np.degrees(x[4]), az / g, ax / g])
