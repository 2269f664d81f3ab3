"""Exact-arithmetic pins for the CPU oracle (tests/golden/exact_*.npz).

The reference ships no golden vectors and cannot be built here (no Rust toolchain), so the oracle's
parity with it stays "unpinned" in the strict sense.  What CAN be pinned: the reference is IEEE f64 arithmetic on
f64 constants, so any faithful f64 build of it is within a few ulp x conditioning of the EXACT value of its formulas
on the same f64 inputs.  This script evaluates those formulas — written here a third time, from the cited Rust lines,
as scalar mpmath code at 60 digits (constants and inputs are the f64 values, converted exactly) — and
tests/test_oracle_cpu.py::test_oracle_against_exact_arithmetic asserts the C oracle is within a stated bound of them:

  * the three MPPI models' dynamics and costs (examples/mppi4.rs:20-27,81-89, mppi4-non-liner.rs:81-94,
    mppi4-non-liner-ukf.rs:33-35,126-148), single steps and H-step rollouts
  * one full Mppi::compute (src/mppi.rs:38-91) per model on seeded noise
  * one predict + update of the Cholesky UKF of examples/ukf-pen.rs:28-141
  * gen_q(dt) (examples/mppi4-non-liner-ukf.rs:192-221), dynamics_short and hx of the same file (:149-179)

    python tests/make_exact.py      # rewrites tests/golden/exact_*.npz (a few seconds)
"""
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 60
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F = mp.mpf


def f(x):
    """exact value of an f64"""
    return F(float(x))


# ---- constants (the f64 literals of the examples, combined exactly) ----
def consts_pen(j2=0.2):
    M1, R_W, L, G, KT = f(150e-3), f(50e-3), f(0.2474), f(9.81), f(0.15)
    M2 = f(2.3) - f(2.0) * M1 + f(2.0)
    return dict(M1=M1, R_W=R_W, M2=M2, L=L, J1=M1 * R_W * R_W, J2=f(j2), G=G, KT=KT)


def consts_nl6():
    return dict(M1=f(160e-3), R_W=f(50e-3), M2=f(2.4), L=f(0.4), J1=f(2.23e5) * f(1e-9), J2=f(1.168e8) * f(1e-9), G=f(9.81), KT=f(0.15))


def clamp(v, lo, hi):
    return lo if v < lo else (hi if v > hi else v)


def cost_clamped(x):  # examples/mppi4.rs:20-27
    xc = clamp(x[0], F(-2), F(2))
    a = clamp(x[1] + 2 * xc, F(-5), F(5))
    b = x[2] + f(0.35) * clamp(x[0], f(-0.75), f(0.75))
    return 2 * xc ** 2 + 3 * a ** 2 + 5 * b ** 2 + f(1.2) * x[3] ** 2


def cost_quadratic(x):  # examples/mppi4-non-liner-ukf.rs:33-35
    return f(0.1) * x[0] ** 2 + f(0.1) * x[1] ** 2 + F(1) * x[2] ** 2 + f(0.5) * x[3] ** 2


def dyn_L(x, u, dt, c):  # examples/mppi4.rs:81-89 (semi-implicit: x3, x2, x1, x0)
    M1, R_W, M2, L, J1, J2, G, KT = (c[k] for k in ("M1", "R_W", "M2", "L", "J1", "J2", "G", "KT"))
    D = (M1 + M2 + J1 / (R_W * R_W)) * (M2 * L * L + J2) - M2 * M2 * L * L
    x = list(x)
    x[3] += ((M1 + M2 + J1 / (R_W * R_W)) / D * M2 * G * L * x[2] - M2 * L / D / R_W * KT * u) * dt
    x[2] += x[3] * dt
    x[1] += (-M2 * M2 * G * L * L / D * x[2] + (M2 * L * L + J2) / D / R_W * KT * u) * dt
    x[0] += x[1] * dt
    return x


def dyn_NL(x, u, dt, c):  # examples/mppi4-non-liner.rs:81-94 (explicit Euler on the old state)
    M1, R_W, M2, L, J1, J2, G, KT = (c[k] for k in ("M1", "R_W", "M2", "L", "J1", "J2", "G", "KT"))
    D = (M1 + M2 + J1 / (R_W * R_W)) * (M2 * L * L + J2)
    s, co = mp.sin(x[2]), mp.cos(x[2])
    d = D - M2 * M2 * L * L * co * co
    term1 = (M1 + M2 + J1 / (R_W * R_W)) * M2 * G * L * s
    q = KT * u / R_W + M2 * L * x[3] ** 2 * s
    term2 = q * M2 * L * co
    term3 = (J2 + M2 * L * L) * q
    term4 = M2 * G * L * L * s * co
    return [x[0] + x[1] * dt, x[1] + (term3 + term4) / d * dt, x[2] + x[3] * dt, x[3] + (term1 - term2) / d * dt]


def ddot6(th, thd, u, fpush, c):  # examples/mppi4-non-liner-ukf.rs:126-139
    M1, R_W, M2, L, J1, J2, G, KT = (c[k] for k in ("M1", "R_W", "M2", "L", "J1", "J2", "G", "KT"))
    D1 = (2 * M1 + M2 + 2 * J1 / (R_W * R_W)) * (M2 * L * L + J2)
    s, co, cf = mp.sin(th), mp.cos(th), mp.cos(thd)
    d = D1 - (M2 * L * co) ** 2
    ddx = ((M2 * L * L + J2) * M2 * L / d * thd ** 2 * s - (M2 * L) ** 2 * G / d * s * co
           + 2 * (M2 * L * L + J2) / (d * R_W) * KT * u + (M2 * L * L + J2) / d * fpush * cf)
    ddth = (-(M2 * L) ** 2 / d * thd ** 2 * s * co + (M2 * G * s - 2 * fpush) * L * (2 * M1 + M2 + 2 * J1 / (R_W * R_W)) / d
            - 2 * M2 * L / (d * R_W) * KT * u * co - M2 * L * fpush * cf ** 2 / d)
    return ddx, ddth


def dyn_NL6(x, u, dt, c):  # dynamics4 :140-148
    ddx, ddth = ddot6(x[2], x[3], u, F(0), c)
    x = list(x)
    x[3] += ddth * dt
    x[2] += x[3] * dt
    x[1] += ddx * dt
    x[0] += x[1] * dt
    return x


def dynamics_short(x6, u, dt, fpush, c):  # :149-159
    ddx, ddth = ddot6(x6[3], x6[4], u, fpush, c)
    r = list(x6)
    r[5] = ddth
    r[4] += r[5] * dt
    r[3] += r[4] * dt
    r[2] = ddx
    r[1] += r[2] * dt
    r[0] += r[1] * dt
    return r


def hx_nl6(x, c):  # :169-179
    G, L, R_W = c["G"], c["L"], c["R_W"]
    ax = G * mp.sin(x[3]) + x[2] * mp.cos(x[3]) + L * x[5]
    az = G * mp.cos(x[3]) - x[2] * mp.sin(x[3]) + L * x[4] ** 2
    return [F(36) * F(60) / (2 * mp.pi * R_W) * x[1], F(36) * F(-60) / (2 * mp.pi * R_W) * x[1], x[4] * 180 / mp.pi, az / G, ax / G]


def gen_q(dt):  # :192-221
    dt2, dt3, dt4 = dt ** 2, dt ** 3, dt ** 4
    q = [[F(0)] * 6 for _ in range(6)]

    def block(i, j, k, w):  # the 3x3 pattern on rows/cols (i, j, k)
        q[i][j] += w * dt4 / 8; q[i][k] += w * dt3 / 6
        q[j][i] += w * dt4 / 8; q[j][j] += w * dt3 / 3; q[j][k] += w * dt2 / 2
        q[k][i] += w * dt3 / 6; q[k][j] += w * dt2 / 2; q[k][k] += w * dt
    block(3, 4, 5, F(100))
    block(1, 3, 4, F(70))
    block(0, 1, 2, F(20))
    return q


MODELS = {0: (dyn_L, cost_clamped, consts_pen), 1: (dyn_NL, cost_clamped, consts_pen), 2: (dyn_NL6, cost_quadratic, consts_nl6)}


def mppi_exact(mid, dt, lam, sig, lim, x, u_n, eps):
    """src/mppi.rs:38-91 in exact arithmetic on f64 inputs; returns (u_new[H], c[K], argmax)."""
    dyn, cost, consts = MODELS[mid]
    c = consts()
    K, H = eps.shape
    dt, lam, inv = f(dt), f(lam), 1 / (f(sig) ** 2)
    lo, hi = f(lim[0]), f(lim[1])
    un = [f(v) for v in u_n]
    vs, cs = [], []
    for k in range(K):
        v = [clamp(un[t] + f(eps[k, t]), lo, hi) for t in range(H)]
        xs, J = [f(a) for a in x], F(0)
        for t in range(H):
            xs = dyn(xs, v[t], dt, c)
            J += cost(xs)
        ctrl = sum((un[t] * inv) * v[t] for t in range(H))
        vs.append(v)
        cs.append(-J - ctrl)
    m = max(cs)
    w = [mp.exp((ck - m) / lam) for ck in cs]
    s = sum(w)
    u_new = [sum((w[k] / s) * vs[k][t] for k in range(K)) for t in range(H)]
    return np.array([float(a) for a in u_new]), np.array([float(a) for a in cs]), int(np.argmax([float(a) for a in cs]))


# ---- Cholesky UKF of examples/ukf-pen.rs:28-141 ----
def ukf_pen_exact(x, P, Q, R, u, z):
    n, o = 4, 2
    c = consts_pen(0.1)
    alpha, beta, kappa = f(1e-3), F(2), F(3 - n)
    C = alpha * alpha * (n + kappa)
    lam_ = C - n
    wm = [lam_ / C] + [1 / (2 * C)] * (2 * n)
    wc = [lam_ / C + 1 - alpha ** 2 + beta] + [1 / (2 * C)] * (2 * n)
    A = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            A[i, j] = C * f(P[i][j])
    Lm = mp.cholesky(A)
    xs = [f(v) for v in x]
    sig = [list(xs)]
    for i in range(n):  # interleaved order (:46-56)
        sig.append([xs[r] + Lm[r, i] for r in range(n)])
        sig.append([xs[r] - Lm[r, i] for r in range(n)])
    sig = [dyn_L(s, f(u), f(0.01), c) for s in sig]

    def ut(points, w_m, w_c, cov, dim):
        mean = [sum(w_m[i] * points[i][r] for i in range(len(points))) for r in range(dim)]
        Pm = [[sum(w_c[i] * (points[i][r] - mean[r]) * (points[i][q] - mean[q]) for i in range(len(points))) + f(cov[r][q])
               for q in range(dim)] for r in range(dim)]
        return mean, Pm
    xp, Pp = ut(sig, wm, wc, Q, n)
    zs = [[s[1], s[3]] for s in sig]  # hx :86-91
    zp, Pz = ut(zs, wm, wc, R, o)
    Pxz = [[sum(wc[i] * (sig[i][r] - xp[r]) * (zs[i][q] - zp[q]) for i in range(len(sig))) for q in range(o)] for r in range(n)]
    Pzi = mp.matrix(Pz) ** -1
    Kg = mp.matrix(Pxz) * Pzi
    innov = mp.matrix([f(z[0]) - zp[0], f(z[1]) - zp[1]])
    xn = mp.matrix(xp) + Kg * innov
    Pn = mp.matrix(Pp) - Kg * mp.matrix(Pz) * Kg.T
    return (np.array([float(xn[i]) for i in range(n)]), np.array([[float(Pn[i, j]) for j in range(n)] for i in range(n)]),
            np.array([[float(a) for a in s] for s in sig]))


def main():
    os.makedirs(GOLD, exist_ok=True)
    rng = np.random.Generator(np.random.PCG64(20240008))
    out = {}
    # models: single steps on random states, and rollouts
    for mid, dt in ((0, 0.1), (1, 0.1), (1, 0.008), (2, 0.15)):
        dyn, cost, consts = MODELS[mid]
        c = consts()
        X = np.concatenate([rng.normal(0, 0.6, (24, 4)), rng.normal(0, 3.0, (8, 4))])
        U = rng.uniform(-20, 20, 32)
        nxt, cst = [], []
        for x, u in zip(X, U):
            r = dyn([f(a) for a in x], f(u), f(dt), c)
            nxt.append([float(a) for a in r])
            cst.append(float(cost(r)))
        tag = f"m{mid}_dt{dt}"
        out[tag + "_x"], out[tag + "_u"], out[tag + "_next"], out[tag + "_cost"] = X, U, np.array(nxt), np.array(cst)
    np.savez_compressed(os.path.join(GOLD, "exact_models.npz"), **out)
    print("wrote exact_models")

    out = {}
    for mid, H, dt, lam, sig, lim, K in ((0, 8, 0.1, 0.5, 3.0, (-20.0, 20.0), 96), (1, 8, 0.1, 0.5, 3.0, (-20.0, 20.0), 96),
                                           (1, 40, 0.02, 0.5, 3.0, (-20.0, 20.0), 48), (2, 8, 0.15, 1.4, 4.0, (-10.0, 10.0), 96)):
        x0 = np.array([0.5, 0.0, 0.1, 0.0])
        u_n = rng.uniform(-2, 2, H)
        eps = sig * rng.standard_normal((K, H))
        u_new, cs, am = mppi_exact(mid, dt, lam, sig, lim, x0, u_n, eps)
        tag = f"m{mid}_H{H}"
        out[tag + "_cfg"] = np.array([mid, H, dt, lam, sig, lim[0], lim[1], K])
        out[tag + "_x"], out[tag + "_u_n"], out[tag + "_eps"], out[tag + "_u_out"], out[tag + "_c"], out[tag + "_argmax"] = x0, u_n, eps, u_new, cs, np.array(am)
    np.savez_compressed(os.path.join(GOLD, "exact_mppi.npz"), **out)
    print("wrote exact_mppi")

    out = {}
    Q = np.zeros((4, 4)); Q[1, 1] = 1.0; Q[2, 2] = 0.25; Q[2, 3] = 0.5; Q[3, 2] = 0.5; Q[3, 3] = 1.0
    R = np.diag([0.5, 0.5])
    xs, Ps, zs, xo, Po = [], [], [], [], []
    for b in range(6):
        x = rng.normal(0, 0.1, 4)
        A = rng.normal(0, 1, (4, 4))
        P = 10.0 * np.eye(4) if b == 0 else (A @ A.T + np.eye(4))
        z = rng.normal(0, 0.7, 2)
        xn, Pn, _ = ukf_pen_exact(x, P, Q, R, 0.0015, z)
        xs.append(x); Ps.append(P); zs.append(z); xo.append(xn); Po.append(Pn)
    out.update(x=np.array(xs), P=np.array(Ps), z=np.array(zs), x_out=np.array(xo), P_out=np.array(Po), Q=Q, R=R, u=np.array(0.0015))
    np.savez_compressed(os.path.join(GOLD, "exact_ukf_pen.npz"), **out)
    print("wrote exact_ukf_pen")

    out = {}
    c6 = consts_nl6()
    for dt in (0.01, 0.0093):
        out[f"gen_q_{dt}"] = np.array([[float(v) for v in row] for row in gen_q(f(dt))])
    X6 = rng.normal(0, 0.4, (16, 6))
    U6 = rng.uniform(-10, 10, 16)
    out["x6"], out["u6"] = X6, U6
    out["short_f0"] = np.array([[float(a) for a in dynamics_short([f(v) for v in x], f(u), f(0.01), F(0), c6)] for x, u in zip(X6, U6)])
    out["short_f2"] = np.array([[float(a) for a in dynamics_short([f(v) for v in x], f(u), f(0.01), F(2), c6)] for x, u in zip(X6, U6)])
    out["hx"] = np.array([[float(a) for a in hx_nl6([f(v) for v in x], c6)] for x in X6])
    np.savez_compressed(os.path.join(GOLD, "exact_nl6.npz"), **out)
    print("wrote exact_nl6")


if __name__ == "__main__":
    main()
