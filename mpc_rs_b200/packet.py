"""Wire packets of the reference's serial link (src/packet.rs): fixed-layout little-endian structs, COBS-framed.

Host-side only — this is the byte format in front of the UKF (Sensor3 -> (enable, z[5]), src/packet.rs:102-120,
consumed by examples/mppi4-ukf-commu.rs:262-263) and behind the MPPI (Control::from_current, :69-76).  Layouts follow
the `#[repr(C)]` / `#[repr(packed)]` declarations (:5-41); zerocopy reads and writes them in native byte order, which
is little-endian on the reference's hosts.  COBS is cobs_rs::stuff / unstuff with marker 0 (:43-61; the crate is not
vendored in the reference, the encoding is the standard consistent-overhead byte stuffing for < 254 payload bytes:
one overhead byte, every 0 replaced by the distance to the next 0, the marker appended).
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass
from typing import ClassVar, Optional, Tuple

import numpy as np


def cobs_stuff(payload: bytes, marker: int = 0) -> bytes:
    """len(payload) + 2 bytes: overhead byte, stuffed payload, marker."""
    if len(payload) >= 254:
        raise ValueError("packets of the reference are < 254 bytes")
    if marker != 0:
        raise ValueError("only marker 0 is used by the reference (src/packet.rs:52,56)")
    out = bytearray(len(payload) + 2)
    code_at = 0
    code = 1
    for i, b in enumerate(payload):
        if b == marker:
            out[code_at] = code
            code_at = i + 1
            code = 1
        else:
            out[i + 1] = b
            code += 1
    out[code_at] = code
    out[-1] = marker
    return bytes(out)


def cobs_unstuff(frame: bytes, marker: int = 0) -> bytes:
    """Inverse of cobs_stuff for a frame of payload length + 2 bytes."""
    if marker != 0:
        raise ValueError("only marker 0 is used by the reference")
    n = len(frame) - 2
    out = bytearray(frame[1:1 + n])
    nxt = frame[0]
    pos = 0
    while nxt != 0 and pos + nxt <= n:
        pos += nxt
        nxt = frame[pos]
        out[pos - 1] = 0
    return bytes(out)


class _Packet:
    FMT: ClassVar[str] = ""

    @classmethod
    def size(cls) -> int:
        return struct.calcsize(cls.FMT)  # SIZE, src/packet.rs:46

    @classmethod
    def buf_size(cls) -> int:
        return cls.size() + 2  # BUF_SIZE, :47

    def _fields(self) -> tuple:
        raise NotImplementedError

    def as_bytes(self) -> bytes:
        return struct.pack(self.FMT, *self._fields())

    def as_cobs(self) -> bytes:  # :50-53
        return cobs_stuff(self.as_bytes(), 0)

    @classmethod
    def from_bytes(cls, raw: bytes):
        raise NotImplementedError

    @classmethod
    def from_cobs(cls, frame: bytes):  # :55-58
        if len(frame) != cls.buf_size():
            return None
        return cls.from_bytes(cobs_unstuff(frame, 0))


@dataclass
class State(_Packet):  # :5-12, 16 bytes
    x: float
    dx: float
    theta: float
    dtheta: float
    FMT: ClassVar[str] = "<4f"

    def _fields(self):
        return (self.x, self.dx, self.theta, self.dtheta)

    @classmethod
    def from_bytes(cls, raw):
        return cls(*struct.unpack(cls.FMT, raw))

    def to_vector(self) -> np.ndarray:  # From<State> for Vector4<f64>, :79-83
        return np.array([self.x, self.dx, self.theta, self.dtheta], dtype=np.float64)


def _as_i16(v: float) -> int:
    """Rust `f64 as i16`: truncate toward zero, saturate, NaN -> 0."""
    if v != v:
        return 0
    return int(max(-32768, min(32767, math.trunc(v)))) if math.isfinite(v) else (32767 if v > 0 else -32768)


@dataclass
class Control(_Packet):  # :14-18, 2 bytes
    u: int
    FMT: ClassVar[str] = "<h"
    MAX: ClassVar[int] = 10000  # :70

    def _fields(self):
        return (self.u,)

    @classmethod
    def from_bytes(cls, raw):
        return cls(*struct.unpack(cls.FMT, raw))

    @classmethod
    def from_current(cls, current: float) -> "Control":  # :71-75: K = MAX / 10, u = (K * current) as i16
        return cls(_as_i16(cls.MAX / 10.0 * float(current)))


@dataclass
class Sensor(_Packet):  # :20-25, 8 bytes
    encoder: Tuple[int, int]
    gyro: float
    FMT: ClassVar[str] = "<2hf"

    def _fields(self):
        return (*self.encoder, self.gyro)

    @classmethod
    def from_bytes(cls, raw):
        e0, e1, g = struct.unpack(cls.FMT, raw)
        return cls((e0, e1), g)

    def to_vector(self) -> np.ndarray:  # :85-89
        return np.array([self.encoder[0], self.encoder[1], self.gyro], dtype=np.float64)


@dataclass
class Sensor2(_Packet):  # :27-33, 16 bytes
    encoder: Tuple[int, int]
    gyro: float
    accel: Tuple[float, float]
    FMT: ClassVar[str] = "<2hf2f"

    def _fields(self):
        return (*self.encoder, self.gyro, *self.accel)

    @classmethod
    def from_bytes(cls, raw):
        e0, e1, g, a0, a1 = struct.unpack(cls.FMT, raw)
        return cls((e0, e1), g, (a0, a1))

    def to_vector(self) -> np.ndarray:  # :91-100
        return np.array([self.encoder[0], self.encoder[1], self.gyro, *self.accel], dtype=np.float64)


@dataclass
class Sensor3(_Packet):  # :35-41, #[repr(packed)]: 17 bytes, no padding after `enable`
    enable: int
    encoder: Tuple[int, int]
    gyro: float
    accel: Tuple[float, float]
    FMT: ClassVar[str] = "<B2hf2f"

    def _fields(self):
        return (self.enable, *self.encoder, self.gyro, *self.accel)

    @classmethod
    def from_bytes(cls, raw):
        en, e0, e1, g, a0, a1 = struct.unpack(cls.FMT, raw)
        return cls(en, (e0, e1), g, (a0, a1))

    def parse(self) -> Tuple[int, np.ndarray]:
        """(enable, z[5]) with the readings of disabled sensors zeroed (:102-120) — z feeds BatchedUkf.update after
        set_enable(enable) and set_r(gen_r(enable, R)) like examples/mppi4-ukf-commu.rs:262-295."""
        z = np.array([self.encoder[0], self.encoder[1], self.gyro, *self.accel], dtype=np.float64)
        for i in range(5):
            if (self.enable & (1 << i)) == 0:
                z[i] = 0.0
        return self.enable, z


def read_frame(stream_bytes: bytes, packet_cls) -> Optional[_Packet]:
    """The `read` helper of examples/mppi4-ukf-commu.rs:243-252: given the bytes up to and including a 0x00
    delimiter, decode the last BUF_SIZE bytes as one packet (None if the buffer is too short)."""
    n = packet_cls.buf_size()
    if len(stream_bytes) < n:
        return None
    return packet_cls.from_cobs(stream_bytes[-n:])
