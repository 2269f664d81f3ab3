"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/mpc_b200.h declares,
argument validation works without a GPU, there is no CPU fallback, and the multi-rank protocol (shard ranges,
id rendezvous, partial-row exchange + merge) is exercised with world_size = 2 over gloo."""
import ctypes as C
import os
import re
import socket

import numpy as np
import pytest

import oracle_lib as O
from mpc_rs_b200 import _abi as A
from mpc_rs_b200 import distributed as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mpc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mpcb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 45
    assert declared == set(A.SYMBOLS), declared ^ set(A.SYMBOLS)  # the Python binding covers the whole header
    lib = C.CDLL(A.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/mpc_b200.h but not exported"
    assert A.lib().mpcb_abi_version() == 1


def test_status_strings_are_the_reference_messages():
    L = A.lib()
    # src/mppi.rs:69,77,88; src/ukf.rs:69; examples/ukf-pen.rs:45
    assert [L.mpcb_status_string(i).decode() for i in range(1, 6)] == [
        "Cannot calculate max", "sum is zero", "u is invalid", "Inverse fail", "Cholesky fail"]


def test_defaults_match_the_examples():
    L = A.lib()
    cfg = A.MppiCfg()
    assert L.mpcb_mppi_default_cfg(A.MODEL_L, C.byref(cfg)) == 0
    # examples/mppi4.rs:8-18
    assert (cfg.horizon, cfg.samples, cfg.lambda_, cfg.std_dev, cfg.limit_lo, cfg.limit_hi) == (8, 800000, 0.5, 3.0, -20.0, 20.0)
    assert cfg.model.dt == pytest.approx(0.1) and cfg.model.m2 == 2.3 - 2.0 * 150e-3 + 2.0 and cfg.precision == A.F32
    assert L.mpcb_mppi_default_cfg(A.MODEL_NL6, C.byref(cfg)) == 0
    # examples/mppi4-non-liner-ukf.rs:13-24; FP64 (fast form) by default (DESIGN.md precision policy)
    assert (cfg.samples, cfg.lambda_, cfg.std_dev, cfg.limit_hi, cfg.precision) == (500000, 1.4, 4.0, 10.0, A.F64_FAST)
    assert list(cfg.model.cost)[:4] == [0.1, 0.1, 1.0, 0.5]
    assert L.mpcb_mppi_default_cfg(A.MODEL_PEN_LIN, C.byref(cfg)) == A.BAD_ARG  # not an MPPI model
    # product defaults == oracle defaults, field by field
    for mid in (A.MODEL_L, A.MODEL_NL, A.MODEL_NL6, A.MODEL_PEN_LIN, A.MODEL_PEN_NL, A.MODEL_PEN6, A.MODEL_NL6_UKF):
        mp = A.ModelParams()
        assert L.mpcb_model_defaults(mid, C.byref(mp)) == 0
        op = O.model_defaults(mid)
        for f, _ in A.ModelParams._fields_[:9]:
            assert getattr(mp, f) == getattr(op, f), (mid, f)
        assert list(mp.cost) == list(op.cost)
    for mid in (A.MODEL_PEN_LIN, A.MODEL_PEN_NL, A.MODEL_PEN6, A.MODEL_NL6_UKF):
        n, o = O.dims(mid)
        Q, R, P0 = np.empty((n, n)), np.empty((o, o)), np.empty((n, n))
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        assert L.mpcb_ukf_default_noise(mid, 0.01, dp(Q), dp(R), dp(P0)) == 0
        Qo, Ro, Po = O.ukf_default_noise(mid, 0.01)
        np.testing.assert_array_equal(Q, Qo)
        np.testing.assert_array_equal(R, Ro)
        np.testing.assert_array_equal(P0, Po)
    ucfg = A.UkfCfg()
    assert L.mpcb_ukf_default_cfg(A.MODEL_PEN_LIN, C.byref(ucfg)) == 0
    assert (ucfg.n, ucfg.o, ucfg.sqrt_mode, ucfg.sigma_order) == (4, 2, A.SQRT_CHOLESKY, A.ORDER_INTERLEAVED)  # examples/ukf-pen.rs
    assert L.mpcb_ukf_default_cfg(A.MODEL_PEN6, C.byref(ucfg)) == 0
    assert (ucfg.n, ucfg.o, ucfg.sqrt_mode, ucfg.sigma_order) == (6, 5, A.SQRT_EIG, A.ORDER_LIBRARY)  # src/ukf2.rs


def test_no_cpu_fallback_and_argument_validation():
    """Without a device every create() fails loudly; bad arguments are BAD_ARG on any box."""
    L = A.lib()
    cfg = A.MppiCfg()
    L.mpcb_mppi_default_cfg(A.MODEL_NL, C.byref(cfg))
    h = A._H()
    bad = A.MppiCfg.from_buffer_copy(cfg)
    bad.horizon = 0
    assert L.mpcb_mppi_create(C.byref(h), C.byref(bad)) == A.BAD_ARG and not h.value
    bad = A.MppiCfg.from_buffer_copy(cfg)
    bad.state_dim = 3
    assert L.mpcb_mppi_create(C.byref(h), C.byref(bad)) == A.BAD_ARG
    bad = A.MppiCfg.from_buffer_copy(cfg)
    bad.rank, bad.world_size = 2, 2
    assert L.mpcb_mppi_create(C.byref(h), C.byref(bad)) == A.BAD_ARG
    assert b"rank" in L.mpcb_last_error_string()
    if L.mpcb_device_count() < 1:
        st = L.mpcb_mppi_create(C.byref(h), C.byref(cfg))
        assert st in (A.CUDA_ERROR, A.BAD_ARG) and not h.value
        from mpc_rs_b200 import Mppi, BatchedUkf, MpcB200Error, models
        with pytest.raises(MpcB200Error):
            Mppi(8, 1000, model=models.L, lam=0.5, std_dev=3.0, limit=(-20, 20))
        with pytest.raises(MpcB200Error):
            BatchedUkf(models.PEN_LIN, 4)
    # the product package never touches the oracle
    import mpc_rs_b200
    pkg = os.path.dirname(mpc_rs_b200.__file__)
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("the oracle", "").replace("an f64 restatement", ""), fn


def test_shard_ranges_tile_exactly():
    for total in (1, 7, 65536, 800000, (1 << 24) + 3):
        for world in (1, 2, 3, 4, 8):
            if total < world:
                continue
            spans = [D.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        D.shard_range(10, 2, 2)


# ---------------------------------------------------------------- world_size = 2 over gloo
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _partial_row(oid, p, H, lam, sig, lim, x, u_n, eps_shard, k_offset):
    """The row one rank contributes: [max, argmax, sum_w, n_finite, sum_w*v[0..H)] (csrc/mppi_kernel.cuh)."""
    K = eps_shard.shape[0]
    st, _, info, c = O.mppi_compute(oid, p, K, H, lam, sig, lim[0], lim[1], x, u_n, eps_shard, want_costs=True)
    v = np.clip(u_n[None, :] + eps_shard, lim[0], lim[1])
    w = np.exp((c - info["max"]) / lam)
    return np.concatenate([[info["max"], float(info["argmax"] + k_offset), w.sum(), float(info["n_finite"])], w @ v])


def _merge_rows(rows, lam):
    M = rows[:, 0].max()
    s = np.exp((rows[:, 0] - M) / lam)
    return (s[:, None] * rows[:, 4:]).sum(0) / (s * rows[:, 2]).sum(), int(rows[np.argmax(rows[:, 0]), 1])


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) the id rendezvous: rank 0's 128 bytes reach everyone
        uid = D.exchange_unique_id(lambda: bytes(range(128)))
        # (1b) the peer-handle rendezvous: everyone ends with every rank's 128 bytes, in rank order
        hs = D.exchange_handles(bytes([rank]) * 128)
        assert hs == [bytes([r]) * 128 for r in range(world)]
        # (2) one sharded control step: each rank reduces its sample shard to one row, one all_gather, same merge
        oid, H, dt, lam, sig, lim, K = O.MODEL_NL, 8, 0.1, 0.5, 3.0, (-20.0, 20.0), 5001
        p = O.model_defaults(oid, dt=dt)
        rng = np.random.default_rng(42)  # same stream on both ranks: the global noise tensor
        eps = sig * rng.standard_normal((K, H))
        x, u_n = np.array([0.5, 0.0, 0.1, 0.0]), rng.uniform(-1, 1, H)
        first, count = D.shard_range(K, rank, world)
        row = torch.from_numpy(_partial_row(oid, p, H, lam, sig, lim, x, u_n, eps[first:first + count], first))
        rows = [torch.empty_like(row) for _ in range(world)]
        dist.all_gather(rows, row)
        u, arg = _merge_rows(torch.stack(rows).numpy(), lam)
        st, u_ref, info, _ = O.mppi_compute(oid, p, K, H, lam, sig, lim[0], lim[1], x, u_n, eps)
        q.put((rank, uid == bytes(range(128)), float(np.linalg.norm(u - u_ref) / np.linalg.norm(u_ref)), arg == info["argmax"]))
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for rank, uid_ok, err, arg_ok in res:
        assert uid_ok and arg_ok
        assert err < 1e-12  # sharding changes only the reduction order


def test_user_model_source_compiles_without_a_gpu():
    """mpcb_mppi_check_user_source: the kernel headers embedded in the library compile under NVRTC around a user's
    dynamics/cost (both precisions), and a broken source comes back as MPCB_RTC_ERROR with the compiler's log."""
    import mpc_rs_b200 as M
    from mpc_rs_b200 import _abi as A
    ok = """
    template <typename real> void dynamics(real (&x)[4], real u, const real* p) { x[0] += x[1] * p[0]; x[1] += u * p[0]; }
    template <typename real> real cost(const real (&x)[4], const real* p) { return x[0] * x[0] + p[1] * (x[1] * x[1]); }
    """
    for prec in ("f32", "f64"):
        assert "error" not in M.check_user_source(ok, prec)
    two = """
    template <typename real> void dynamics(real (&x)[2], real u, const real* p) { x[0] += x[1] * p[0]; x[1] += u * p[0]; }
    template <typename real> real cost(const real (&x)[2], const real* p) { return x[0] * x[0]; }
    """
    assert "error" not in M.check_user_source(two, "f32", S=2)  # Mppi<N,K,S> is generic in S
    import pytest
    with pytest.raises(M.MpcB200Error) as e:
        M.check_user_source("float cost(const float (&x)[4], const float* p) { return nope; }")
    assert e.value.status == A.RTC_ERROR and "nope" in str(e.value) and "user_model.cu(1)" in str(e.value)
    assert A.status_string(A.RTC_ERROR) == "user model did not compile"
    # the UKF twin: fx / hx for n = 3, o = 2
    ukf_src = """
    void fx(double (&x)[3], double u, double dt, const double* p) { x[0] += x[1] * dt; x[1] += (u - p[0] * sin(x[0])) * dt; x[2] *= 0.9; }
    void hx(const double (&x)[3], double (&z)[2], const double* p) { z[0] = x[0]; z[1] = x[1] + x[2]; }
    """
    assert "error" not in M.check_user_ukf_source(ukf_src, 3, 2)
    with pytest.raises(M.MpcB200Error) as e:
        M.check_user_ukf_source(ukf_src, 4, 2)  # the signatures say N = 3
    assert e.value.status == A.RTC_ERROR
