"""Wire packets (src/packet.rs): struct sizes and layouts, COBS known answers, Control::from_current, Sensor3::parse."""
import struct

import numpy as np

from mpc_rs_b200 import packet as P


def test_sizes_match_the_rust_layouts():
    # std::mem::size_of of the #[repr(C)] / #[repr(packed)] structs (src/packet.rs:5-41); BUF_SIZE = SIZE + 2 (:47)
    assert (P.State.size(), P.Control.size(), P.Sensor.size(), P.Sensor2.size(), P.Sensor3.size()) == (16, 2, 8, 16, 17)
    assert P.Sensor3.buf_size() == 19 and P.Control.buf_size() == 4


def test_cobs_known_answers():
    # the worked examples of the COBS paper / reference implementations, marker 0
    kat = [(b"\x00", b"\x01\x01\x00"), (b"\x00\x00", b"\x01\x01\x01\x00"), (b"\x11\x22\x00\x33", b"\x03\x11\x22\x02\x33\x00"),
           (b"\x11\x22\x33\x44", b"\x05\x11\x22\x33\x44\x00"), (b"\x11\x00\x00\x00", b"\x02\x11\x01\x01\x01\x00")]
    for raw, enc in kat:
        assert P.cobs_stuff(raw) == enc
        assert P.cobs_unstuff(enc) == raw
    rng = np.random.default_rng(0)
    for _ in range(200):
        raw = bytes(rng.integers(0, 4, rng.integers(1, 40), dtype=np.uint8))  # many zeros
        enc = P.cobs_stuff(raw)
        assert len(enc) == len(raw) + 2 and 0 not in enc[:-1] and enc[-1] == 0
        assert P.cobs_unstuff(enc) == raw


def test_control_from_current_is_a_saturating_truncating_cast():
    # K = 10000 / 10 = 1000; `as i16` truncates toward zero and saturates (src/packet.rs:69-76)
    cases = {0.0: 0, 1.0: 1000, -2.5: -2500, 0.0009: 0, -0.0009: 0, 9.9999: 9999, 40.0: 32767, -40.0: -32768,
             float("nan"): 0, float("inf"): 32767}
    for cur, want in cases.items():
        assert P.Control.from_current(cur).u == want, cur
    c = P.Control.from_current(-1.234)
    assert c.as_bytes() == struct.pack("<h", -1234)
    assert P.Control.from_cobs(c.as_cobs()) == c


def test_sensor3_roundtrip_and_parse():
    s = P.Sensor3(0b10101, (1200, -1200), -3.5, (0.98, 0.02))
    raw = s.as_bytes()
    assert len(raw) == 17 and raw[0] == 0b10101 and raw[1:3] == struct.pack("<h", 1200)  # packed: no padding byte
    frame = s.as_cobs()
    assert len(frame) == 19 and frame[-1] == 0
    back = P.Sensor3.from_cobs(frame)
    assert back.enable == s.enable and back.encoder == s.encoder
    assert np.float32(back.gyro) == np.float32(-3.5) and np.float32(back.accel[0]) == np.float32(0.98)
    enable, z = back.parse()
    assert enable == 0b10101
    np.testing.assert_array_equal(z, [1200.0, 0.0, -3.5, 0.0, np.float64(np.float32(0.02))])  # disabled readings -> 0
    assert P.Sensor3.from_cobs(frame[:-1]) is None  # wrong length
    # the stream reader takes the last BUF_SIZE bytes before the delimiter (examples/mppi4-ukf-commu.rs:243-252)
    assert P.read_frame(b"\x07\x07" + frame, P.Sensor3).encoder == (1200, -1200)
    assert P.read_frame(frame[:5], P.Sensor3) is None


def test_other_packets_roundtrip():
    st = P.State(0.5, 0.0, 0.125, -0.25)  # f32-representable values
    assert P.State.from_cobs(st.as_cobs()) == st and st.to_vector().dtype == np.float64
    se = P.Sensor((10, -10), 1.5)
    assert P.Sensor.from_cobs(se.as_cobs()) == se
    np.testing.assert_array_equal(se.to_vector(), [10.0, -10.0, 1.5])
    s2 = P.Sensor2((1, 2), 0.5, (1.0, 0.25))
    assert P.Sensor2.from_cobs(s2.as_cobs()) == s2
    np.testing.assert_array_equal(s2.to_vector(), [1.0, 2.0, 0.5, 1.0, 0.25])
