"""The C++ host mirror (include/mpc_b200.hpp) and the C++ example drivers: they must compile against the C ABI on a
box without a GPU, and on the B200 the mirror's parity test (tests/cpp/mirror_test.cpp, oracle-checked) and the
examples must run."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "mpc_rs_b200")
BUILD = os.path.join(ROOT, "build", "cpp")
FLAGS = ["-std=c++17", "-O2", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include")]
LINK = ["-L" + LIBDIR, "-lmpc_b200", "-Wl,-rpath," + LIBDIR]


def _build(src, out, extra=()):
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, out)
    cmd = ["g++", *FLAGS, os.path.join(ROOT, src), *extra, *LINK, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def _build_all():
    ora = ["-I" + os.path.join(ROOT, "oracle"), "-L" + os.path.join(ROOT, "oracle"), "-lmpc_oracle",
           "-Wl,-rpath," + os.path.join(ROOT, "oracle")]
    return (_build("tests/cpp/mirror_test.cpp", "mirror_test", ora), _build("examples/cpp/mppi4.cpp", "mppi4"),
            _build("examples/cpp/ukf_pen2.cpp", "ukf_pen2"))


def test_cpp_mirror_and_examples_compile():
    """No GPU needed: the header instantiates (Mppi<8,800000,4>, both UKFs, Gaussian) and links against the C ABI."""
    for exe in _build_all():
        assert os.path.exists(exe)


@pytest.mark.gpu
def test_cpp_mirror_parity_and_examples_run(gpu_required, tmp_path):
    mirror, mppi4, ukf_pen2 = _build_all()
    r = subprocess.run([mirror], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "mirror ok" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
    # examples/mppi4.rs through C++: 2 s of closed loop, same printed line and CSV record, pendulum inside 60 degrees
    csv = str(tmp_path / "mppi.csv")
    r = subprocess.run([mppi4, "2.0", csv], capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("t: ")]
    assert len(lines) == 20 and "over 60 degrees" not in r.stdout and "elapsed:" in r.stdout
    data = np.loadtxt(csv, dtype="float", delimiter=",")
    assert data.shape == (20, 6) and abs((data[1, 0] - data[0, 0]) - 0.1) < 1e-12 and np.all(np.abs(data[:, 4]) < np.radians(60))
    r = subprocess.run([ukf_pen2, "30", "5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and len(r.stdout.splitlines()) == 30 and "nan" not in r.stdout.lower(), r.stdout[-2000:] + r.stderr[-2000:]
