"""Host-side pieces of the example drivers (no GPU): the plants restated in mpc_rs_b200/plants.py against the oracle,
the CSV log format of the reference (examples/mppi4.rs:56-63, examples/mppi4-non-liner-ukf.rs:365-386), and the
sensor gating (gen_r / masked hx, examples/mppi4-ukf-commu.rs:228-236,279-293) on the oracle side."""
import numpy as np

import oracle_lib as O
from mpc_rs_b200 import csvlog
from mpc_rs_b200.plants import PlantL, PlantNL, PlantNL6, PlantPen6, PlantPenLin, PlantPenNL


def test_plants_match_the_oracle():
    rng = np.random.default_rng(0)
    for cls, oid, dt in ((PlantL, O.MODEL_L, 0.1), (PlantNL, O.MODEL_NL, 0.1), (PlantNL, O.MODEL_NL, 0.008)):
        plant, p = cls(dt), O.model_defaults(oid, dt=dt)
        for _ in range(20):
            x, u = rng.normal(0, 0.5, 4), rng.uniform(-20, 20)
            np.testing.assert_allclose(plant.step(x, u), O.dynamics(oid, p, x, u), rtol=1e-13, atol=1e-15)
    # batch axis
    xs, us = rng.normal(0, 0.5, (7, 4)), rng.uniform(-5, 5, 7)
    out = PlantNL(0.1).step(xs, us)
    for i in range(7):
        np.testing.assert_allclose(out[i], PlantNL(0.1).step(xs[i], us[i]), rtol=1e-15)
    pl, p = PlantPenLin(), O.model_defaults(O.MODEL_PEN_LIN)
    x = rng.normal(0, 0.3, 4)
    np.testing.assert_allclose(pl.fx(x, 0.0015), O.fx(O.MODEL_PEN_LIN, p, x, 0.0015), rtol=1e-13, atol=1e-16)
    # truth models of the library-UKF examples (examples/ukf-pen2.rs:31-53, examples/ukf-pen3.rs:35-63)
    for cls, oid, n in ((PlantPenNL, O.MODEL_PEN_NL, 4), (PlantPen6, O.MODEL_PEN6, 6)):
        pl, p = cls(), O.model_defaults(oid)
        assert pl.dt == 0.01
        for _ in range(10):
            x = rng.normal(0, 0.4, n)
            np.testing.assert_allclose(pl.fx(x, 0.1), O.fx(oid, p, x, 0.1), rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(pl.hx(x), O.hx(oid, p, x), rtol=1e-12, atol=1e-12)
    p6, pn = PlantNL6(), O.model_defaults(O.MODEL_NL6_UKF)
    for f in (0.0, 2.0):
        x6 = rng.normal(0, 0.3, 6)
        np.testing.assert_allclose(p6.dynamics_short(x6, 1.7, 0.01, f), O.dynamics_short(pn, x6, 1.7, 0.01, f), rtol=1e-12,
                                   atol=1e-14)
    np.testing.assert_allclose(p6.hx(x6), O.hx(O.MODEL_NL6_UKF, pn, x6), rtol=1e-13, atol=1e-13)
    assert p6.push(0.5) == 0.0 and p6.push(1.2) == 2.0 and p6.push(1.5) == 0.0


def test_csv_rows_look_like_rust_to_string(tmp_path):
    # Rust's f64::to_string: shortest round-trip digits, no exponent, no trailing ".0"
    assert [csvlog.fmt(v) for v in (0.1, 100.0, -2.5, 1e-7, 1.5e20, 0.0)] == \
        ["0.1", "100", "-2.5", "0.0000001", "150000000000000000000", "0"]
    path = tmp_path / "logs" / "mppi" / "mppi.csv"
    with csvlog.MppiLog(str(path)) as log:
        log.write(0.0, 1.25, [0.5, 0.0, 0.1, 0.0])
        log.write(0.1, -20.0, [0.5, 1e-3, 0.1, -0.25])
    data = np.loadtxt(str(path), dtype="float", delimiter=",")  # what scripts/plot-mppi.py does (:27-31)
    assert data.shape == (2, 6) and data[1, 1] == -20.0 and data[1, 3] == 1e-3
    assert abs((data[1, 0] - data[0, 0]) - 0.1) < 1e-15  # DT = data_set[1, 0] - data_set[0, 0] (:33)
    path2 = tmp_path / "u.csv"
    with csvlog.MppiUkfLog(str(path2)) as log:
        log.write(0.03, 0.5, np.arange(6), np.arange(6) + 0.5, np.arange(6) - 0.5)
    assert np.loadtxt(str(path2), delimiter=",").shape == (20,)
    try:
        with csvlog.MppiLog(str(tmp_path / "bad.csv")) as log:
            log._row([1, 2, 3])
        raise AssertionError("a short row must be rejected")
    except ValueError:
        pass


def test_sensor_gating_on_the_oracle():
    oid, p = O.MODEL_NL6_UKF, O.model_defaults(O.MODEL_NL6_UKF)
    Q, R, P0 = O.ukf_default_noise(oid, 0.01)
    Rg = O.gen_r(R, 0b10101)
    assert np.array_equal(np.diag(Rg), [200.0, 1e6, 10.0, 1e6, 0.05])  # examples/mppi4-ukf-commu.rs:228-236
    assert np.array_equal(O.gen_r(R, 0b11111), R)
    rng = np.random.default_rng(3)
    B = 4
    x, P = rng.normal(0, 0.05, (B, 6)), np.tile(P0 * 0.01, (B, 1, 1))
    z = rng.normal(0, 1, (B, 5))
    a = O.ukf_step_batch(oid, p, x, P, Q, R, 0.3, z, 0.01, O.SQRT_EIG, O.ORDER_LIBRARY)
    b = O.ukf_step_batch(oid, p, x, P, Q, R, 0.3, z, 0.01, O.SQRT_EIG, O.ORDER_LIBRARY, enable=0xFFFFFFFF)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])  # all sensors on == the plain update
    # a disabled sensor: its hx row reads 0 and (with gen_r) its reading barely moves the estimate
    z2 = z.copy()
    z2[:, 1] += 1000.0
    m1 = O.ukf_step_batch(oid, p, x, P, Q, Rg, 0.3, z, 0.01, O.SQRT_EIG, O.ORDER_LIBRARY, enable=0b10101)
    m2 = O.ukf_step_batch(oid, p, x, P, Q, Rg, 0.3, z2, 0.01, O.SQRT_EIG, O.ORDER_LIBRARY, enable=0b10101)
    assert np.max(np.abs(m1[0] - m2[0])) < 0.05 * np.max(np.abs(a[0] - m1[0]) + 1e-3) or np.max(np.abs(m1[0] - m2[0])) < 1e-2
    assert not np.array_equal(a[0], m1[0])
