"""BASELINE config #4 and SURVEY.md 8(f): sensor gating, the batched MPPI+UKF closed loop against the oracle running
the same schedule, and the example drivers (examples/*.py) end to end."""
import os
import sys

import numpy as np
import pytest

import oracle_lib as O
from mpc_rs_b200 import BatchedUkf, models, ukf
from mpc_rs_b200.closed_loop import ClosedLoopBatch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("name,oid", [("NL6_UKF", O.MODEL_NL6_UKF), ("PEN6", O.MODEL_PEN6)])
def test_sensor_gating_parity(gpu_required, name, oid):
    """set_enable + gen_r (examples/mppi4-ukf-commu.rs:228-236,279-293) against the masked oracle update."""
    model = getattr(models, name)
    p = O.model_defaults(oid)
    Q, R, P0 = O.ukf_default_noise(oid, 0.01)
    rng = np.random.default_rng(7)
    B = 96
    x0, z = rng.normal(0, 0.02, (B, 6)), rng.normal(0, 1.0, (B, 5))
    for enable in (0b11111, 0b10101, 0b00001, 0b11110):
        with BatchedUkf(model, B, exact=True) as f:
            f.init(np.zeros(6), P0 * 0.01, Q, R)
            f.set_state(x0, None)
            Rg = f.gen_r(enable, R)
            assert np.array_equal(Rg, O.gen_r(R, enable))
            f.set_r(Rg)
            f.set_enable(enable)
            f.step(0.4, z, dt=0.01)
            xg, Pg = f.get_state()
        xo, Po, st = O.ukf_step_batch(oid, p, x0, np.tile(P0 * 0.01, (B, 1, 1)), Q, Rg, 0.4, z, 0.01, O.SQRT_EIG,
                                      O.ORDER_LIBRARY, enable=enable)
        assert not st.any()
        assert rel(xg, xo) < 1e-8 and rel(Pg, Po) < 1e-8, (name, bin(enable), rel(xg, xo), rel(Pg, Po))


def _oracle_tick(loop, pm, pu, Q, R, x_true, x_est, P_est, u_seq, u0, k, eps, z=None, noise=None):
    """The CPU restatement of one tick (same order as csrc/closed_loop.cu): returns the new (x_true, z, x_est, P_est, u_seq, u0)."""
    C, H, dt = loop.C, loop.H, loop.tick_dt
    f = loop.plant.push(k * dt)
    x_true = np.stack([O.dynamics_short(pu, x_true[c], u0[c], dt, f) for c in range(C)])
    if z is None:
        z = np.stack([O.hx(O.MODEL_NL6_UKF, pu, x_true[c]) for c in range(C)]) + noise * PlantR
    x_est, P_est, st = O.ukf_step_batch(O.MODEL_NL6_UKF, pu, x_est, P_est, Q, R, u0, z, dt, O.SQRT_EIG, O.ORDER_LIBRARY)
    assert not st.any()
    u_seq = u_seq.copy()
    for c in range(C):
        s, u_new, info, _ = O.mppi_compute(O.MODEL_NL6, pm, loop.K, H, loop.LAMBDA, loop.R_U, loop.LIMIT[0], loop.LIMIT[1],
                                           x_est[c][[0, 1, 3, 4]], u_seq[c], eps[c])
        u_seq[c] = u_new if s == 0 else 0.0
    return x_true, z, x_est, P_est, u_seq, u_seq[:, 0].copy()


def test_closed_loop_matches_the_oracle_schedule(gpu_required):
    """C robots through mpcb_closed_loop_tick_replay (MPPI noise and sensor readings supplied): every tick the GPU loop and
    an oracle loop (same plant, same readings, same noise) must agree on the true state, the applied control, the
    control sequence and the estimate."""
    C, K, ticks = 3, 2048, 5
    rng = np.random.default_rng(99)
    x0 = np.zeros((C, 6))
    x0[:, 3] = [0.05, -0.08, 0.02]
    pm, pu = O.model_defaults(O.MODEL_NL6), O.model_defaults(O.MODEL_NL6_UKF)
    with ClosedLoopBatch(C, K, precision="f64", exact_ukf=True, x0=x0, seed=5) as loop:
        H, dt = loop.H, loop.tick_dt
        Q, R, P0 = O.ukf_default_noise(O.MODEL_NL6_UKF, dt)
        x_true, x_est, P_est = x0.copy(), x0.copy(), np.tile(P0, (C, 1, 1))
        u_seq, u0 = np.zeros((C, H)), np.zeros(C)
        for k in range(ticks):
            eps = loop.R_U * rng.standard_normal((C, K, H))
            x_true, z, x_est, P_est, u_seq, u0 = _oracle_tick(loop, pm, pu, Q, R, x_true, x_est, P_est, u_seq, u0, k, eps,
                                                              noise=rng.normal(0, 1, (C, 5)))
            # GPU side, fed the same sensor readings and noise
            u0_g = loop.tick(eps=eps, z=z)
            xg, Pg = loop.estimate()
            assert np.allclose(loop.x, x_true, rtol=1e-12, atol=1e-14)  # device plant == oracle plant
            # per-tick parity of the coupled parts: UKF 1e-6 on this ill-conditioned n = 6 filter (the sigma weights
            # amplify rounding 1.7e5 x per step), MPPI 1e-6 on the estimate it was given
            assert rel(xg, x_est) < 1e-7 and rel(Pg, P_est) < 1e-6, (k, rel(xg, x_est), rel(Pg, P_est))
            assert rel(loop.controls(), u_seq) < 1e-6, (k, rel(loop.controls(), u_seq))
            assert rel(u0_g, u0) < 1e-6
            # the filter multiplies any difference ~200 x per tick, so both sides continue from the oracle's state:
            # every tick is then a fresh per-step comparison of the whole coupled tick
            loop.set_truth(x_true)
            loop.set_estimate(x_est, P_est)
            loop.set_controls(u_seq)
            assert not loop.mppi_status().any()


def test_closed_loop_device_sensor_and_sharding(gpu_required):
    """Generate mode (device plant, device sensor noise, Philox MPPI noise): the readings the device sensor produced are
    the oracle's hx of the device's true state plus noise of the right scale, and a batch sharded over two handles by
    controller_offset reproduces the unsharded batch (the noise counters use the global robot index: same noise, bit for bit)."""
    C, K, ticks = 8, 1024, 3
    rng = np.random.default_rng(3)
    x0 = np.zeros((C, 6))
    x0[:, 3] = rng.uniform(-0.1, 0.1, C)
    pu = O.model_defaults(O.MODEL_NL6_UKF)
    with ClosedLoopBatch(C, K, x0=x0, seed=11) as whole, ClosedLoopBatch(3, K, x0=x0[:3], seed=11) as lo, \
            ClosedLoopBatch(5, K, x0=x0[3:], seed=11, controller_offset=3) as hi:
        for k in range(ticks):
            for loop in (whole, lo, hi):
                loop.tick()
            xw, zw = whole.x, whole.readings()
            hx = np.stack([O.hx(O.MODEL_NL6_UKF, pu, xw[c]) for c in range(C)])
            noise = (zw - hx) / PlantR
            assert np.all(np.abs(noise) < 6.0) and np.any(np.abs(noise) > 0.05), noise
            # the three handles split the samples over blocks differently (chunks per controller = SMs / C), so the FP64 sums
            # of a control differ at rounding level; everything upstream of the first control is bit-identical
            if k == 0:
                assert np.array_equal(np.vstack([lo.x, hi.x]), xw) and np.array_equal(np.vstack([lo.readings(), hi.readings()]), zw)
            # ... and the loop feeds a rounding-level difference of a control back through the filter, which multiplies it several hundred times per tick
            tol = 1e-10 * 1000.0 ** k
            assert rel(np.vstack([lo.x, hi.x]), xw) < tol and rel(np.vstack([lo.readings(), hi.readings()]), zw) < tol
            assert rel(np.vstack([lo.controls(), hi.controls()]), whole.controls()) < 100.0 * tol
        l0 = whole.launches
        whole.tick()
        whole.sync()
        assert whole.launches - l0 == 5  # plant+sensor, UKF, gather, MPPI, first control — and no host round trip


def test_closed_loop_config4_subset_parity(gpu_required):
    """BASELINE config #4 at its named size — C = 4096 robots x K = 8192 samples — against the oracle on a subset: the first
    64 robots of the big batch and a 64-robot batch (same seed, same initial states) walk the same trajectories (their
    kernels split the samples differently, so sums differ at rounding level), and the 64-robot batch's tick, replayed
    with the noise it drew, matches the oracle schedule."""
    Cbig, Csub, K, ticks = 4096, 64, 8192, 3
    rng = np.random.default_rng(20240004)
    x0 = np.zeros((Cbig, 6))
    x0[:, 3] = rng.uniform(-0.1, 0.1, Cbig)
    pm, pu = O.model_defaults(O.MODEL_NL6), O.model_defaults(O.MODEL_NL6_UKF)
    with ClosedLoopBatch(Cbig, K, x0=x0, seed=7, precision="f64", exact_ukf=True) as big, \
            ClosedLoopBatch(Csub, K, x0=x0[:Csub], seed=7, precision="f64", exact_ukf=True) as sub:
        for k in range(ticks):
            big.tick()
            sub.tick()
            xb, xs = big.x[:Csub], sub.x
            ub, us = big.controls()[:Csub], sub.controls()
            eb, es = big.estimate()[0][:Csub], sub.estimate()[0]
            assert not big.mppi_status().any() and not sub.mppi_status().any()
            if k == 0:  # same noise, same readings: everything upstream of the first control is bit-identical
                assert np.array_equal(xb, xs) and np.array_equal(eb, es)
            # the two batches split the samples over blocks differently: rounding-level differences of the controls, which the
            # loop (plant -> sensor -> filter) multiplies several hundred times per tick
            tol = 1e-10 * 1000.0 ** k
            # (an MPPI step itself turns a 1e-10 difference of its input state into ~1e-7 of the controls on this model)
            assert rel(xb, xs) < tol and rel(eb, es) < tol and rel(ub, us) < 100.0 * tol, (k, rel(xb, xs), rel(eb, es), rel(ub, us))
        # one more tick of the subset against the oracle, with the noise the MPPI kernel draws (dumped) and the device's readings
        H, dt = sub.H, sub.tick_dt
        Q, R, P0 = O.ukf_default_noise(O.MODEL_NL6_UKF, dt)
        x_true, (x_est, P_est), u_seq, u0 = sub.x, sub.estimate(), sub.controls(), sub.applied()
        eps = sub.R_U * rng.standard_normal((Csub, K, H))
        z = np.stack([O.hx(O.MODEL_NL6_UKF, pu, O.dynamics_short(pu, x_true[c], u0[c], dt, sub.plant.push(sub.t))) for c in range(Csub)])
        z = z + rng.normal(0, 1, (Csub, 5)) * PlantR
        xt, z, xe, Pe, useq, u0n = _oracle_tick(sub, pm, pu, Q, R, x_true, x_est, P_est, u_seq, u0, sub.ticks, eps, z=z)
        u0_g = sub.tick(eps=eps, z=z)
        assert np.allclose(sub.x, xt, rtol=1e-12, atol=1e-14)
        assert rel(sub.estimate()[0], xe) < 1e-6 and rel(sub.controls(), useq) < 1e-6 and rel(u0_g, u0n) < 1e-6


PlantR = np.array([200.0, 200.0, 10.0, 0.05, 0.05])


def test_closed_loop_keeps_the_robots_upright(gpu_required):
    """Trajectory-level sanity for the batch: with the estimate in the loop, robots started within +-0.1 rad stay
    inside the reference's abort bound |theta| <= pi/2 through the 2 N push window start."""
    import mppi4_non_liner_ukf as ex
    up = ex.run(controllers=64, samples=4096, seconds=1.2, quiet=True, csv=os.path.join(ROOT, "gpurun_out", "cl_test.csv"))
    assert up.sum() >= 60, f"only {int(up.sum())} of 64 robots upright"
    data = np.loadtxt(os.path.join(ROOT, "gpurun_out", "cl_test.csv"), delimiter=",")
    assert data.ndim == 2 and data.shape[1] == 20 and np.all(np.diff(data[:, 0]) > 0)


def test_example_mppi4_balances_and_logs(gpu_required, tmp_path):
    """examples/mppi4.rs / mppi4-non-liner.rs drivers: 3 s of closed loop, pendulum stays inside 60 degrees, CSV loads
    the way scripts/plot-mppi.py loads it."""
    import mppi4 as ex
    for nonlinear in (False, True):
        path = str(tmp_path / f"mppi_{int(nonlinear)}.csv")
        rows = ex.run(nonlinear=nonlinear, samples=200_000, seconds=3.0, csv=path, quiet=True, seed=11)
        assert len(rows) == 30 and np.all(np.abs(rows[:, 4]) < np.radians(60.0))
        assert abs(rows[-1, 4]) < abs(rows[0, 4]) + 0.05  # theta is being driven back
        data = np.loadtxt(path, dtype="float", delimiter=",")
        assert data.shape == (30, 6) and np.allclose(data, rows, rtol=0, atol=0)


def test_example_ukf_pen_converges(gpu_required):
    import ukf_pen as ex
    x_act, x_est, p = ex.run(batch=256, steps=100, quiet=True, seed=4)
    assert np.all(np.isfinite(x_est)) and np.all(np.isfinite(p))
    # observed components (x1, x3) are tracked to within the sensor noise; P has contracted from 10 I
    assert np.sqrt(np.mean((x_est[:, 1] - x_act[:, 1]) ** 2)) < 0.5
    assert np.all(p[:, 1, 1] < 5.0) and np.all(p[:, 3, 3] < 5.0)


@pytest.mark.parametrize("six", [False, True])
def test_example_library_ukf_runs_like_the_reference(gpu_required, six, capsys):
    """examples/ukf-pen2.rs / ukf-pen3.rs drivers (API-usage fixtures): 100 open-loop steps through
    new / predict / update / state / covariance; the filter stays finite, P stays symmetric with a positive diagonal,
    the observed velocity is tracked, and the printed line has the reference's fields."""
    import ukf_pen2 as ex
    x_act, x_est, p = ex.run(six=six, steps=100, quiet=False, seed=4)
    n = 6 if six else 4
    assert x_est.shape == (n,) and p.shape == (n, n)
    assert np.all(np.isfinite(x_est)) and np.all(np.isfinite(p))
    assert np.allclose(p, p.T, rtol=1e-9, atol=1e-12) and np.all(np.diag(p) > 0)
    assert np.diag(p)[1] < 10.0          # the wheel odometry observes x' (P0 = 10 I)
    lines = capsys.readouterr().out.strip().splitlines()
    assert len(lines) == 100 and lines[0].startswith("t: 0.00 x_act: (") and " x_obs: (" in lines[0] and " p: (" in lines[0]


def test_example_mppi4_non_liner_s(gpu_required, tmp_path):
    """examples/mppi4-non-liner-s.rs driver: MPPI fed by the 4-state library UKF with run-time dt on the fixed schedule;
    the loop runs, estimates stay finite, the observed wheel speed is tracked (theta is not observed by this filter, so
    its estimate is only as good as the model, like in the reference), and the CSV has the reference's 6 columns."""
    import mppi4_non_liner_s as ex
    csv = str(tmp_path / "mppi.csv")
    rows = ex.run(samples=65536, seconds=1.5, csv=csv, quiet=True, seed=2)
    assert rows.shape[1] == 10 and len(rows) >= 50
    assert np.all(np.isfinite(rows))
    assert np.all(np.abs(rows[:, 1]) <= 10.0)                      # LIMIT
    assert np.median(np.abs(rows[:, 3] - rows[:, 7])) < 0.5         # x' (wheel odometry, 50 rpm noise ~ 0.26 m/s)
    data = np.loadtxt(csv, delimiter=",")                          # write_record only: no header row (:157-166)
    assert data.shape == (len(rows), 6)
    np.testing.assert_allclose(data[:, 1], rows[:, 1], rtol=0, atol=0)


def test_handles_are_send_and_distinct_handles_run_concurrently(gpu_required):
    """The reference's examples run the controller in the main thread while a second thread owns the filter
    (examples/mppi4-non-liner-ukf.rs:162-168,273-284: Arc<Mutex<UnscentedKalmanFilter>>): a handle must be usable from
    a thread other than the one that created it (Send), and distinct handles must work concurrently.  Two worker
    threads drive an MPPI handle and a UKF handle created by the main thread at the same time (ctypes drops the GIL
    inside the calls); the results must be the bits a single-threaded run produces."""
    import threading
    from mpc_rs_b200 import Mppi
    H, K, B, T = 16, 20000, 3000, 25
    rng = np.random.default_rng(12)
    eps = (3.0 * rng.standard_normal((T, K, H))).astype(np.float32)
    x0 = np.array([0.5, 0.0, 0.1, 0.0])
    Q, R, P0 = ukf.default_noise(models.PEN_LIN)
    zs = rng.normal(0, 0.7, (T, B, 2))

    def run_mppi(m, out):
        u = np.zeros(H)
        for t in range(T):
            u = m.compute_replay(x0, u, eps[t])
            out.append(u.copy())

    def run_ukf(f, out):
        for t in range(T):
            f.predict(0.0015)
            f.update(zs[t])
        out.extend(f.get_state())

    kw = dict(model=models.NL, lam=0.5, std_dev=3.0, limit=(-20, 20), precision="f32", dt=0.05)
    with Mppi(H, K, **kw) as m1, Mppi(H, K, **kw) as m2, BatchedUkf(models.PEN_LIN, B) as f1, BatchedUkf(models.PEN_LIN, B) as f2:
        for f in (f1, f2):
            f.init(np.zeros(4), P0, Q, R)
        ref_m, ref_f, thr_m, thr_f, errors = [], [], [], [], []
        run_mppi(m1, ref_m)  # single-threaded reference
        run_ukf(f1, ref_f)

        def guard(fn, *a):
            try:
                fn(*a)
            except Exception as e:  # noqa: BLE001 - surfaced below
                errors.append(e)

        ta = threading.Thread(target=guard, args=(run_mppi, m2, thr_m))
        tb = threading.Thread(target=guard, args=(run_ukf, f2, thr_f))
        ta.start(); tb.start(); ta.join(); tb.join()
        assert not errors, errors
        assert len(thr_m) == T and all(np.array_equal(a, b) for a, b in zip(ref_m, thr_m))
        assert np.array_equal(ref_f[0], thr_f[0]) and np.array_equal(ref_f[1], thr_f[1])
