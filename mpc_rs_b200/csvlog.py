"""CSV logs in the formats the reference's examples write and scripts/plot-mppi.py reads (np.loadtxt, no header).

    mppi:      t, u, x0, x1, x2, x3                                   examples/mppi4.rs:56-63
    mppi+ukf:  t, u, x[0..6), x_est[0..6), x_pred[0..6)  (20 columns)   examples/mppi4-non-liner-ukf.rs:365-386

Numbers are printed like Rust's f64::to_string: shortest digits that round-trip, never an exponent.
"""
from __future__ import annotations

import os

import numpy as np


def fmt(v: float) -> str:
    v = float(v)
    if v != v:
        return "NaN"
    if v in (float("inf"), float("-inf")):
        return "inf" if v > 0 else "-inf"
    s = np.format_float_positional(v, trim="-")
    return "-0" if s == "-0" else s


class CsvLog:
    def __init__(self, path: str, columns: int):
        d = os.path.dirname(path)
        if d:
            os.makedirs(d, exist_ok=True)
        self.columns = columns
        self._f = open(path, "w", newline="")

    def _row(self, vals):
        if len(vals) != self.columns:
            raise ValueError(f"expected {self.columns} columns, got {len(vals)}")
        self._f.write(",".join(fmt(v) for v in vals) + "\n")
        self._f.flush()  # the reference flushes every record (examples/mppi4.rs:64)

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class MppiLog(CsvLog):
    """logs/mppi/mppi.csv of examples/mppi4.rs:35-36."""

    def __init__(self, path: str = "logs/mppi/mppi.csv"):
        super().__init__(path, 6)

    def write(self, t, u, x):
        self._row([t, u, *np.asarray(x, dtype=np.float64).reshape(4)])


class MppiUkfLog(CsvLog):
    """The 20-column record of examples/mppi4-non-liner-ukf.rs:357-386."""

    def __init__(self, path: str = "logs/mppi/mppi.csv"):
        super().__init__(path, 20)

    def write(self, t, u, x, x_est, x_pred):
        self._row([t, u, *np.asarray(x, dtype=np.float64).reshape(6), *np.asarray(x_est, dtype=np.float64).reshape(6),
                   *np.asarray(x_pred, dtype=np.float64).reshape(6)])
