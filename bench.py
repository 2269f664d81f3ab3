#!/usr/bin/env python
"""bench.py — MPPI rollout-steps/s (and batched-UKF filter-updates/s) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" is one MPPI control step (src/mppi.rs:33-92: noise -> K x H rollout -> softmax-weighted update) on
BASELINE.json configs[1]: examples/mppi4-non-liner.rs model NL, K = 65536 samples, H = 100 (DT = 0.8/100), FP32
fast path, Philox noise generated in-register.  With N > 1 (torchrun, one rank per GPU) every rank rolls out its
own 65536-sample shard of a K = 65536*N controller (weak scaling); the ranks' partial sums meet INSIDE the rollout kernel
(tagged 16-byte cells written into every peer's mailbox over NVLink; MPCB_BENCH_TRANSPORT=nccl selects one ncclAllGather
per control step instead).

    value     device-resident closed loop (u_out of step i is u_in of step i+1), CUDA events per step on the
              handle's stream, L2 flushed between timed steps, max over ranks
    e2e       the same step through the C-ABI call mpcb_mppi_compute with HOST buffers (inputs travel in the kernel
              parameters, u_out/info come back through mapped pinned memory as self-validating cells the host polls), host wall clock
              around a compiled closed loop (tools/e2e_loop.c, like the reference's compiled callers); the same loop
              written in Python is reported next to it (python_loop_value)
    roofline  algorithmic FP32 flops (60 per rollout-step) / kernel time vs the FFMA peak measured
              live by tools/peak_bench (MEASURED_PEAKS.json has no FP32 vector number)
    ukf       BASELINE configs[2] on the side: 2^20 independent examples/ukf-pen.rs filters per GPU, FP64,
              336 algorithmic bytes per filter-update vs the measured HBM copy bandwidth; its own cpu_baseline (the
              oracle looped over filters with OpenMP on all host cores)
    closed_loop  BASELINE configs[3]: 4096 robots x 8192 samples x H = 8 of examples/mppi4-non-liner-ukf.rs through
              mpcb_closed_loop_tick (plant, sensor, UKF, MPPI all on the device), robots sharded over the N ranks
    sweep     BASELINE configs[4]: K = 2^16 .. 2^24, H = 200, samples sharded over the N ranks
    parity_check  (N > 1) untimed: the sharded controller against a single-GPU controller of the same K and seed
    cpu_baseline / --impl reference
              the C restatement of the reference's nalgebra/rayon CPU path (oracle/, kind "port": no Rust
              toolchain exists here or on the GPU box) on all host cores
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ---- workload: BASELINE.json configs[1] ------------------------------------------------------------------
K_PER_GPU = 65536
H = 100
DT = 0.8 / H  # T = 0.8 s as shipped (examples/mppi4-non-liner.rs:8), N raised to 100
LAMBDA, SIGMA, LIMIT = 0.5, 3.0, (-20.0, 20.0)
X0 = np.array([0.5, 0.0, 0.1, 0.0])  # examples/mppi4-non-liner.rs:30
FLOPS_PER_STEP = 60.0  # model NL 58 + 2 epilogue flops per rollout-step
UKF_B = 1 << 20
UKF_T = 100  # T = 100 steps, as examples/ukf-pen.rs:154-178 runs
UKF_BYTES = 336.0  # 8*(2n + 2n^2 + o), n=4, o=2
# dram__bytes_read.sum + dram__bytes_write.sum per launch from this round's ncu --set full captures of these kernels
# (profiles/mppi_r2e_ncu_full_summary.txt, profiles/ukf_r2e_ncu_full_summary.txt; tools/prof_round.sh) — the one roofline
# field that cannot be measured inside this run
MPPI_DRAM_TRAFFIC_BYTES = 70_912 + 0
UKF_DRAM_TRAFFIC_BYTES = 138_444_288 + 62_852_096
FP32_FALLBACK_TFLOPS = 69.5  # tools/peak_bench on this pool's B200 (profiles/peaks_r1.json)
FP64_FALLBACK_TFLOPS = 33.9
METRIC = "mppi_rollout_steps_per_sec"
UNIT = "rollout-steps/s"


def workload(n_gpus: int) -> dict:
    transport = os.environ.get("MPCB_BENCH_TRANSPORT", "peer")
    exchange = (f"one ncclAllGather of {H + 4} doubles per rank per step" if transport == "nccl" else
                "in-kernel peer exchange (no collective call): every merger warp stores its column pair's sums into all ranks' "
                "mailboxes over NVLink as self-validating tagged 16-byte cells and combines the ranks' cells itself")
    return {
        "workload": "BASELINE configs[1]: examples/mppi4-non-liner.rs MPPI, model NL, "
                    f"K={K_PER_GPU} samples/GPU x H={H}, DT={DT}, lambda={LAMBDA}, sigma={SIGMA}, limit=+-20",
        "samples_per_gpu": K_PER_GPU, "samples_total": K_PER_GPU * n_gpus, "horizon": H, "controllers": 1,
        "noise": "philox4x32-7 in-register (generate mode)", "precision": "f32 rollout, f64 accumulation/softmax",
        "sharding": f"samples x{n_gpus}, {exchange}" if n_gpus > 1 else "none",
        "l2": "flushed between timed steps (256 MiB write); inputs are O(H) bytes",
    }


# ---- reference arm / cpu baseline ---------------------------------------------------------------------------
def cpu_baseline(steps, warmup: int, budget_s: float = 12.0, n_gpus: int = 1):
    """Times the CPU restatement (oracle/) of src/mppi.rs:38-91 on all host cores: full-size control steps of the bench
    workload (K = 65536, H = 100), closed loop.  steps=None runs as many as fit in about budget_s seconds of CPU work
    (the bounded sample); an explicit step count is capped by the same budget.  Returns (rollout-steps/s, info)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every core this process may run on
    threads = max(O.max_threads(), len(os.sched_getaffinity(0)))
    p = O.model_defaults(O.MODEL_NL, dt=DT)
    K = K_PER_GPU * n_gpus  # the GPU arm's whole-job workload (weak scaling: K_PER_GPU samples per GPU)
    u = np.zeros(H)
    t0 = time.perf_counter()
    O.mppi_compute_cpu(O.MODEL_NL, p, K, H, LAMBDA, SIGMA, LIMIT[0], LIMIT[1], X0, u, seed=1, threads=threads)
    one = time.perf_counter() - t0
    fit = max(3, int(budget_s / max(one, 1e-6)))
    steps = fit if steps is None else max(1, min(int(steps), fit))
    for i in range(warmup):
        O.mppi_compute_cpu(O.MODEL_NL, p, K, H, LAMBDA, SIGMA, LIMIT[0], LIMIT[1], X0, u, seed=2 + i, threads=threads)
    x, t0 = X0.copy(), time.perf_counter()
    for i in range(steps):
        st, u, _ = O.mppi_compute_cpu(O.MODEL_NL, p, K, H, LAMBDA, SIGMA, LIMIT[0], LIMIT[1], x, u, seed=100 + i,
                                      threads=threads)
    el = time.perf_counter() - t0
    return K * H * steps / el, {
        "kind": "port", "cores": threads, "steps": steps,
        "sample": f"{steps} control steps ({el:.1f} s of wall time on {threads} threads) of K={K} x H={H} (model NL) through "
                  "oracle/ orc_mppi_compute_cpu: materialised v[K][H], six passes, xoshiro256+/ziggurat per worker, f64, "
                  "-O3 -march=native -ffp-contract=off; C restatement of the nalgebra/rayon path, not the Rust binary",
        "ms_per_step": el / steps * 1e3,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    val, info = cpu_baseline(args.steps, args.warmup, budget_s=120.0, n_gpus=args.gpus)  # every step is a full-size control step
    steps = info["steps"]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": info["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload(args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---- helpers ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons every ~2 ms through NVML in a background thread while the timed
    regions run (nvidia-smi -lms cannot sample a region that lasts a few milliseconds)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop, self._thr = False, None

    def start(self):
        try:
            import threading
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}

            def loop():
                while not self._stop:
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for nm, bit in bits.items():
                            if r & bit:
                                self.reasons.add(nm)
                    except Exception:
                        pass
                    time.sleep(0.002)

            self._thr = threading.Thread(target=loop, daemon=True)
            self._thr.start()
        except Exception:
            self._thr = None

    def stop(self) -> dict:
        self._stop = True
        if self._thr is not None:
            self._thr.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peaks():
    """FP32 / FP64 vector peaks from tools/peak_bench (run live), HBM from MEASURED_PEAKS.json."""
    peaks = {"fp32_tflops": FP32_FALLBACK_TFLOPS, "fp64_tflops": FP64_FALLBACK_TFLOPS, "fp32_source": "fallback (profiles/peaks_r1.json)",
             "hbm_gbs": 6650.0, "hbm_source": "fallback (B200_PROFILING.md)"}
    exe = os.path.join(ROOT, "tools", "peak_bench")
    if os.path.exists(exe):
        try:
            r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
            j = json.loads(r.stdout.strip().splitlines()[-1])
            if j.get("err") == "no error" and j["fp32_tflops"] > 1:
                peaks.update(fp32_tflops=j["fp32_tflops"], fp64_tflops=j["fp64_tflops"],
                             fp32_source="measured live by tools/peak_bench (FFMA/DFMA chains, ILP 8)")
        except Exception:
            pass
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peaks.update(hbm_gbs=float(mp["hbm_gbs"]), hbm_source="measured (MEASURED_PEAKS.json)")
    except Exception:
        pass
    return peaks


def dev_alloc(A, dev, nbytes):
    p = C.c_void_p()
    A.check(A.lib().mpcb_device_alloc(dev, nbytes, C.byref(p)))
    return p.value


def upload(A, dev, dptr, arr):
    arr = np.ascontiguousarray(arr)
    A.check(A.lib().mpcb_device_upload(dev, dptr, arr.ctypes.data_as(C.c_void_p), arr.nbytes))



# ---- BASELINE configs[3]: the batched closed loop, robots sharded over the ranks (no exchange) ----------------------
CL_C, CL_K, CL_TICKS = 4096, 8192, 50


def closed_loop_leg(dev, rank, world, barrier, max_over_ranks):
    from mpc_rs_b200.closed_loop import ClosedLoopBatch
    lo, hi = CL_C * rank // world, CL_C * (rank + 1) // world
    rng = np.random.default_rng(20240004)
    x0 = np.zeros((CL_C, 6))
    x0[:, 3] = rng.uniform(-0.1, 0.1, CL_C)  # theta0 ~ U(-0.1, 0.1)
    out = {"config": {"workload": f"BASELINE configs[3]: examples/mppi4-non-liner-ukf.rs closed loop, C={CL_C} robots (sharded x{world}) x "
                                  f"K={CL_K} samples x H=8 (model NL6 + UKF NL6_UKF), fixed tick 0.01 s, MPPI fed the UKF estimate, "
                                  f"{CL_TICKS} ticks through mpcb_closed_loop_tick", "controllers": CL_C, "samples": CL_K, "horizon": 8}}
    for prec in ("f64fast", "f64", "f32"):
        with ClosedLoopBatch(hi - lo, CL_K, x0=x0[lo:hi], seed=20240004, precision=prec, device=dev, controller_offset=lo) as loop:
            loop.tick(5)
            loop.sync()
            barrier()
            l0, t0 = loop.launches, time.perf_counter()
            loop.tick(CL_TICKS)
            loop.sync()
            el = max_over_ranks(time.perf_counter() - t0)
            up = int(loop.upright().sum())
            bad = int((loop.mppi_status() != 0).sum())
            out[prec] = {"value": CL_C * CL_K * 8 * CL_TICKS / el, "unit": UNIT, "ms_per_tick": el / CL_TICKS * 1e3,
                         "filter_updates_per_sec": CL_C * CL_TICKS / el, "gpu_launches": int(loop.launches - l0),
                         "upright_on_rank0": f"{up}/{hi - lo}", "mppi_failures_on_rank0": bad}
    out["f64fast"]["note"] = ("the default precision for model NL6 (MPCB_F64_FAST): every operation FP64, the FP32 kernels' folded "
                              "formulas with one reciprocal per step and FMA; controls within 1e-9 of the f64 reference, same argmin "
                              "(tests/test_mppi_gpu.py::test_replay_parity_f64_fast)")
    out["f64fast"]["roofline"] = {"bound": "fp64", "flops_per_rollout_step": 72.0,
                                  "achieved": out["f64fast"]["value"] / world * 72.0 / 1e12, "unit": "TFLOP/s",
                                  "note": "72 algorithmic flops per NL6 rollout-step (SURVEY.md 8d), FP64; whole tick incl. plant, sensor and UKF"}
    out["f64"]["note"] = "MPCB_F64: the reference's operation order without FMA contraction (bit-level twin of the f64 oracle up to libm)"
    out["f32"]["note"] = ("FP32 rollouts: on model NL6 at its shipped DT = 0.15 the FP32 controls are 1e-4..1e-3 from the f64 "
                          "reference (the model is chaotic inside its horizon, DESIGN.md 4.1) - reported, not the parity path")
    out["f32"]["roofline"] = {"bound": "fp32", "flops_per_rollout_step": 72.0,
                              "achieved": out["f32"]["value"] / world * 72.0 / 1e12, "unit": "TFLOP/s",
                              "note": "72 algorithmic flops per NL6 rollout-step (SURVEY.md 8d); whole tick incl. plant, sensor and UKF"}
    return out


# ---- BASELINE configs[4]: K = 2^16 .. 2^24, H = 200, samples sharded over the ranks ----------------------------------
def sweep_leg(dev, rank, world, barrier, max_over_ranks, A, Mppi, models, torch, dev_loop):
    from mpc_rs_b200.distributed import attach_mppi_peers
    Hs, rows = 200, []
    for lg in (16, 18, 20, 22, 24):
        K = 1 << lg
        m = Mppi(Hs, K, model=models.NL, lam=LAMBDA, std_dev=SIGMA, limit=LIMIT, precision="f32", dt=0.8 / Hs, device=dev, rank=rank,
                 world_size=world, seed=20240005)
        if world > 1:
            attach_mppi_peers(m)
        d_x, d_a, d_b = dev_alloc(A, dev, 32), dev_alloc(A, dev, 8 * Hs), dev_alloc(A, dev, 8 * Hs)
        upload(A, dev, d_x, X0)
        upload(A, dev, d_a, np.zeros(Hs))
        stream = torch.cuda.ExternalStream(m.stream, device=dev)
        steps = 20 if lg <= 20 else 8

        def run(n):
            if dev_loop is not None:
                if dev_loop(m._h, d_x, d_a, d_b, n) != 0:
                    raise SystemExit("sweep: mpcb_mppi_compute_device failed")
            else:
                for i in range(n):
                    m.compute_device(d_x, d_a if i % 2 == 0 else d_b, d_b if i % 2 == 0 else d_a)
        run(4)
        m.sync()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run(steps)
        e1.record(stream)
        m.sync()
        ms = max_over_ranks(e0.elapsed_time(e1) / steps)
        info = m.last_info()[0]
        rows.append({"K": K, "ms_per_step": ms, "value": K * Hs / (ms * 1e-3), "status": int(info["status"]),
                     "tflops_fp32_algorithmic_per_gpu": K * Hs / (ms * 1e-3) / world * FLOPS_PER_STEP / 1e12})
        for q in (d_x, d_a, d_b):
            A.lib().mpcb_device_free(dev, q)
        m.close()
    return {"config": {"workload": f"BASELINE configs[4]: MPPI sweep K=2^16..2^24 (global), H={Hs}, model NL, DT={0.8 / Hs}, FP32, generate "
                                   f"mode, samples sharded x{world}" + (", in-kernel peer exchange" if world > 1 else ""), "horizon": Hs},
            "unit": UNIT, "points": rows}

# ---- GPU arm --------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from mpc_rs_b200 import BatchedUkf, Mppi, models
    from mpc_rs_b200 import _abi as A

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N > 1 with torchrun (one rank per GPU); see the module docstring")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libmpc_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- controller: global K = K_PER_GPU * world, this rank's shard = K_PER_GPU samples ----
    mppi = Mppi(H, K_PER_GPU * world, model=models.NL, lam=LAMBDA, std_dev=SIGMA, limit=LIMIT, precision="f32", dt=DT,
                device=dev, rank=rank, world_size=world, seed=20240001)
    transport = os.environ.get("MPCB_BENCH_TRANSPORT", "peer")  # "peer": fused in-kernel exchange; "nccl": ncclAllGather
    if world > 1:
        from mpc_rs_b200.distributed import attach_mppi, attach_mppi_peers
        if transport != "nccl":
            # every rank's mailbox handle -> everyone, then cudaIpcOpenMemHandle; if any rank cannot map its peers
            # (e.g. devices hidden from each other) all ranks fall back to the NCCL transport together
            ok = 1
            try:
                attach_mppi_peers(mppi)
            except Exception as e:  # noqa: BLE001 - reported below
                ok = 0
                print(f"[bench] rank {rank}: peer exchange unavailable ({e}); falling back to NCCL", file=sys.stderr)
            t_ok = torch.tensor([ok], dtype=torch.int32, device="cuda")
            dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
            if int(t_ok.item()) == 0:
                transport = "nccl"
                os.environ["MPCB_BENCH_TRANSPORT"] = "nccl"
                mppi.close()
                mppi = Mppi(H, K_PER_GPU * world, model=models.NL, lam=LAMBDA, std_dev=SIGMA, limit=LIMIT, precision="f32", dt=DT,
                            device=dev, rank=rank, world_size=world, seed=20240001)
        if transport == "nccl":
            attach_mppi(mppi)  # rank 0's ncclUniqueId -> everyone (torch.distributed broadcast), then ncclCommInitRank

    stream = torch.cuda.ExternalStream(mppi.stream, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    d_x = dev_alloc(A, dev, 32)
    d_u = [dev_alloc(A, dev, 8 * H), dev_alloc(A, dev, 8 * H)]
    upload(A, dev, d_x, X0)
    upload(A, dev, d_u[0], np.zeros(H))

    def device_step(i):
        mppi.compute_device(d_x, d_u[i & 1], d_u[(i + 1) & 1])

    for i in range(args.warmup):
        device_step(i)
    mppi.sync()

    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = mppi.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    with torch.cuda.stream(stream):
        for i in range(args.steps):
            flush.zero_()  # L2 flush (256 MiB write) on the handle's stream, outside the event pair
            evs[i][0].record(stream)
            device_step(args.warmup + i)
            evs[i][1].record(stream)
    mppi.sync()
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = max_over_ranks(float(np.sum(step_ms)))
    launches = mppi.launches - launches0
    steps_total = K_PER_GPU * world * H * args.steps
    value = steps_total / (total_ms * 1e-3)
    info = mppi.last_info()[0]
    if info["status"] != 0:
        raise SystemExit(f"MPPI reported status {info['status']} in the timed region")

    # back-to-back variant (no flush, launches queued): kernel duration proper for the roofline
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the launches are enqueued from compiled code (tools/e2e_loop.c: mpcb_device_loop, ~3 us of host time per step); a
    # Python loop over the same call takes ~30 us per step, as long as the kernel, and would time the interpreter
    dev_loop = None
    e2e_so = os.path.join(ROOT, "tools", "libmpcb_e2e.so")
    if os.path.exists(e2e_so):
        dev_loop = C.CDLL(e2e_so).mpcb_device_loop
        dev_loop.restype = C.c_int
        dev_loop.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    n_kern = max(args.steps, 100)
    e0.record(stream)
    if dev_loop is not None:
        st_k = dev_loop(mppi._h, d_x, d_u[0], d_u[1], n_kern)
        if st_k != 0:
            raise SystemExit(f"mpcb_mppi_compute_device returned {st_k} in the back-to-back loop")
    else:
        for i in range(n_kern):
            device_step(i)
    e1.record(stream)
    mppi.sync()
    kern_ms = max_over_ranks(e0.elapsed_time(e1) / n_kern)

    # ---- e2e: host buffers through the C-ABI entry point mpcb_mppi_compute itself (what a Rust/C++/ctypes caller of
    # include/mpc_b200.h calls): x[4], u_in[H] -> u_out[H] + info, closed loop (u_out is the next u_in) ----
    x_h, u_h, out_h = X0.copy(), np.zeros(H), np.zeros(H)
    info_h = (A.MppiInfo * 1)()
    px, pu, po = (a.ctypes.data_as(C.POINTER(C.c_double)) for a in (x_h, u_h, out_h))
    compute = A.lib().mpcb_mppi_compute

    def e2e_step():
        st = compute(mppi._h, px, pu, po, info_h)
        if st != 0:
            raise SystemExit(f"mpcb_mppi_compute returned {st} in the e2e loop")
        np.copyto(u_h, out_h)

    # the loop itself runs compiled (tools/e2e_loop.c -> tools/libmpcb_e2e.so): the reference's callers are compiled
    # programs looping over Mppi::compute (examples/mppi4.rs:41-68).  The same loop written in Python is timed too.
    e2e_lib = None
    e2e_so = os.path.join(ROOT, "tools", "libmpcb_e2e.so")
    if os.path.exists(e2e_so):
        e2e_lib = C.CDLL(e2e_so)
        e2e_lib.mpcb_e2e_loop.restype = C.c_double
        e2e_lib.mpcb_e2e_loop.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                          C.c_int, C.c_int, C.POINTER(C.c_int)]
    for _ in range(args.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    e2e_py_s = max_over_ranks(time.perf_counter() - t0)
    e2e_s, e2e_caller = e2e_py_s, "python loop over the ctypes call"
    if e2e_lib is not None:
        st_c = C.c_int(0)
        barrier()
        el = e2e_lib.mpcb_e2e_loop(mppi._h, px, pu, po, H, args.steps, C.byref(st_c))
        if el < 0:
            raise SystemExit(f"mpcb_mppi_compute returned {st_c.value} in the compiled e2e loop")
        e2e_s, e2e_caller = max_over_ranks(el), "compiled loop (tools/e2e_loop.c) over mpcb_mppi_compute"
    e2e_value = steps_total / e2e_s
    clocks = sampler.stop() if rank == 0 else None

    # ---- UKF, BASELINE configs[2]: 2^20 examples/ukf-pen.rs filters per GPU ----
    ukf_out = None
    try:
        f = BatchedUkf(models.PEN_LIN, UKF_B, device=dev)
        Q, R, P0 = __import__("mpc_rs_b200").ukf.default_noise(models.PEN_LIN)
        f.init(np.zeros(4), P0, Q, R)
        g = torch.Generator(device="cuda").manual_seed(20240003 + rank)
        z = 0.7 * torch.randn((UKF_T, 2, UKF_B), dtype=torch.float64, device="cuda", generator=g)
        ustream = torch.cuda.ExternalStream(f.stream, device=dev)
        torch.cuda.synchronize()
        zs = z.element_size() * 2 * UKF_B
        for t in range(3):
            f.run_device(1, z.data_ptr() + t * zs, u=0.0015)
        f.sync()
        barrier()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ul0 = f.launches
        u0.record(ustream)
        for t in range(UKF_T):  # one launch per filter step: x, P, z in; x, P out (336 B per update)
            f.run_device(1, z.data_ptr() + t * zs, u=0.0015)
        u1.record(ustream)
        f.sync()
        step_ms_ukf = max_over_ranks(u0.elapsed_time(u1))
        f.init(np.zeros(4), P0, Q, R)
        barrier()
        u0.record(ustream)
        f.run_device(UKF_T, z.data_ptr(), u=0.0015)  # T steps fused: state stays in registers
        u1.record(ustream)
        f.sync()
        fused_ms_ukf = max_over_ranks(u0.elapsed_time(u1))
        st = f.status()
        ukf_launches = f.launches - ul0
        upd = UKF_B * UKF_T * world
        ukf_out = {
            "metric": "ukf_filter_updates_per_sec", "unit": "filter-updates/s",
            "value": upd / (step_ms_ukf * 1e-3), "fused_value": upd / (fused_ms_ukf * 1e-3),
            "config": {"workload": f"BASELINE configs[2]: examples/ukf-pen.rs UKF (n=4,o=2, Cholesky), B={UKF_B} filters/GPU, "
                                   f"{UKF_T} steps, FP64, SoA", "filters_per_gpu": UKF_B, "steps": UKF_T},
            "ms_per_step": step_ms_ukf / UKF_T, "fused_ms_per_step": fused_ms_ukf / UKF_T, "dtype": "f64",
            "failed_filters": int((st != 0).sum()), "gpu_launches": int(ukf_launches),
        }
        f.close()
        del z
    except Exception as e:  # the MPPI line is still valid; say what happened
        ukf_out = {"error": repr(e)}

    # ---- UKF CPU baseline (rank 0, N = 1): the oracle's examples/ukf-pen.rs step looped over filters with OpenMP ----
    if world == 1 and ukf_out and "value" in ukf_out:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib as O
            threads = max(O.max_threads(), len(os.sched_getaffinity(0)))
            Bc, Tc = 1 << 18, 41  # ~10 s of CPU work
            pu = O.model_defaults(O.MODEL_PEN_LIN)
            Qc, Rc, P0c = O.ukf_default_noise(O.MODEL_PEN_LIN, 0.0)
            xc, Pc = np.zeros((Bc, 4)), np.tile(P0c, (Bc, 1, 1))
            zc = 0.7 * np.random.default_rng(3).standard_normal((Tc, Bc, 2))
            xc, Pc, _ = O.ukf_step_batch(O.MODEL_PEN_LIN, pu, xc, Pc, Qc, Rc, 0.0015, zc[0], 0.0, O.SQRT_CHOLESKY, O.ORDER_INTERLEAVED,
                                         threads=threads)
            t0 = time.perf_counter()
            for t in range(1, Tc):
                xc, Pc, _ = O.ukf_step_batch(O.MODEL_PEN_LIN, pu, xc, Pc, Qc, Rc, 0.0015, zc[t], 0.0, O.SQRT_CHOLESKY,
                                             O.ORDER_INTERLEAVED, threads=threads)
            el_c = time.perf_counter() - t0
            ukf_out["cpu_baseline"] = {
                "value": Bc * (Tc - 1) / el_c, "unit": "filter-updates/s", "cores": threads, "kind": "port",
                "sample": f"{Tc - 1} steps of {Bc} filters ({el_c:.1f} s on {threads} threads) through oracle/ orc_ukf_step_batch "
                          "(examples/ukf-pen.rs:93-141 per filter, OpenMP over filters, f64, incl. the ctypes array copies); the "
                          "reference itself runs one filter on one thread"}
        except Exception as e:  # noqa: BLE001
            ukf_out["cpu_baseline"] = {"error": repr(e)}

    # ---- BASELINE configs[3] and configs[4] as sub-objects (every rank takes part) ----
    try:
        closed_loop = closed_loop_leg(dev, rank, world, barrier, max_over_ranks)
    except Exception as e:  # noqa: BLE001
        closed_loop = {"error": repr(e)}
    try:
        sweep = sweep_leg(dev, rank, world, barrier, max_over_ranks, A, Mppi, models, torch, dev_loop)
    except Exception as e:  # noqa: BLE001
        sweep = {"error": repr(e)}

    # ---- N > 1: correctness of the sharded path, untimed — a fresh sharded controller against a fresh single-GPU one
    # with the same K and seed (same Philox counters => same samples), and the ranks' results against each other ----
    parity = None
    if world > 1:
        try:
            from mpc_rs_b200.distributed import attach_mppi, attach_mppi_peers
            kw = dict(model=models.NL, lam=LAMBDA, std_dev=SIGMA, limit=LIMIT, precision="f32", dt=DT, device=dev, seed=20240001)
            sh = Mppi(H, K_PER_GPU * world, rank=rank, world_size=world, **kw)
            (attach_mppi if transport == "nccl" else attach_mppi_peers)(sh)
            u_test = np.linspace(-1.0, 1.0, H)
            upload(A, dev, d_u[0], u_test)
            sh.compute_device(d_x, d_u[0], d_u[1])
            sh.sync()
            u_sh = np.empty(H)
            A.check(A.lib().mpcb_device_download(dev, u_sh.ctypes.data_as(C.c_void_p), d_u[1], 8 * H))
            arg_sh = sh.last_info()[0]["argmax"]
            t_all = [torch.zeros(H, dtype=torch.float64, device="cuda") for _ in range(world)]
            dist.all_gather(t_all, torch.from_numpy(u_sh).cuda())
            bitwise = all(bool(torch.equal(t_all[0], t)) for t in t_all)
            sh.close()
            if rank == 0:
                with Mppi(H, K_PER_GPU * world, **kw) as one:
                    u_one = one.compute(X0, u_test)
                    arg_one = one.last_call_info()[0]["argmax"]
                parity = {"rel_err": float(np.linalg.norm(u_sh - u_one) / np.linalg.norm(u_one)), "argmax_equal": bool(arg_sh == arg_one),
                          "ranks_bitwise_equal": bool(bitwise),
                          "what": f"one untimed step: K={K_PER_GPU * world} sharded x{world} vs the same K on one GPU (same seed: the "
                                  "Philox counter is the global sample index)"}
        except Exception as e:  # noqa: BLE001
            parity = {"error": repr(e)}

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = measured_peaks()
    if isinstance(sweep, dict):
        for row in sweep.get("points", []):  # the same FP32 roofline as the headline kernel, per sweep point
            row["frac_fp32_roofline_per_gpu"] = row["tflops_fp32_algorithmic_per_gpu"] / peaks["fp32_tflops"]
    ach_tflops = K_PER_GPU * H * FLOPS_PER_STEP / (kern_ms * 1e-3) / 1e12
    roof = {
        "bound": "fp32", "kernel": "mppi_ws_kernel<ModelNL, 7 packed consumer warps + 9 producer warps, generate> (csrc/mppi_ws_kernel.cuh)",
        "achieved": ach_tflops, "peak": peaks["fp32_tflops"], "unit": "TFLOP/s", "frac": ach_tflops / peaks["fp32_tflops"],
        "traffic": MPPI_DRAM_TRAFFIC_BYTES, "kernel_ms": kern_ms, "kernel_ms_with_l2_flush": total_ms / args.steps,
        "note": f"{FLOPS_PER_STEP:.0f} algorithmic FP32 flops per rollout-step (SURVEY.md 8d) x {K_PER_GPU * H} steps per launch; "
                f"peak = {peaks['fp32_source']}; neither 'hbm' nor 'tensor' bounds this kernel: it is bound by the SM's instruction "
                "dispatch (66 issued instructions per rollout-step; this instruction mix caps at frac 0.36, and at 0.22 on this shape "
                "with its fixed launch + merge tail, DESIGN.md 4.1b) and moves 71 KB of DRAM per launch (traffic: bytes, "
                "this round's ncu capture); kernel_ms = launch-to-launch time of back-to-back launches enqueued from compiled code, no L2 flush; the launches carry "
                "programmatic stream serialization, so the next grid is scheduled while the previous one drains (MPCB_PDL=0: plain launches, "
                "+2.5 us); kernel_ms_with_l2_flush is the same step behind a 256 MiB fill, where nothing overlaps"
                + ("; at N > 1 it includes the in-kernel cross-GPU exchange" if world > 1 else ""),
    }
    if ukf_out and "value" in ukf_out:
        gbs = ukf_out["value"] / world * UKF_BYTES / 1e9
        ukf_out["roofline"] = {"bound": "hbm", "kernel": "ukf_kernel<4,2,PEN_LIN,cholesky,interleaved,fused>", "achieved": gbs,
                               "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "traffic": UKF_DRAM_TRAFFIC_BYTES,
                               "traffic_gbs": UKF_DRAM_TRAFFIC_BYTES / (ukf_out["ms_per_step"] * 1e-3) / 1e9,
                               "traffic_frac": UKF_DRAM_TRAFFIC_BYTES / (ukf_out["ms_per_step"] * 1e-3) / 1e9 / peaks["hbm_gbs"],
                               "note": f"{UKF_BYTES:.0f} algorithmic bytes per filter-update (x, full P in and out, z: SURVEY.md 8d); the "
                                       "kernel neither reads nor writes the strictly-upper triangle of the exactly symmetric P, so the DRAM "
                                       "traffic is lower than the algorithmic figure: traffic_frac is the fraction of the HBM peak in bytes "
                                       f"actually moved (the kernel sits between the HBM and the FP64 roof, DESIGN.md 4.2); peak = {peaks['hbm_source']}; "
                                       f"FP64 pipe peak {peaks['fp64_tflops']:.1f} TFLOP/s ({peaks['fp32_source']})"}
    cpu_val, cpu_info = cpu_baseline(None, 2, budget_s=12.0) if world == 1 else (None, None)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload(world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * (4 + H) * world,
                "d2h_bytes_per_step": (8 * H + 40) * world, "ms_per_step": e2e_s / args.steps * 1e3, "caller": e2e_caller,
                "python_loop_value": steps_total / e2e_py_s, "python_loop_ms_per_step": e2e_py_s / args.steps * 1e3},
        "gpu_launches": int(launches), "roofline": roof, "clocks": clocks,
        "state_updates_per_sec": value * 4, "ukf": ukf_out, "closed_loop": closed_loop, "sweep": sweep,
    }
    if parity is not None:
        line["parity_check"] = parity
    if cpu_val is not None:
        line["cpu_baseline"] = {"value": cpu_val, "unit": UNIT, "cores": cpu_info["cores"], "kind": cpu_info["kind"],
                                "sample": cpu_info["sample"]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
