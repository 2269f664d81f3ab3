"""BASELINE config #5: MPPI scaling sweep K = 2^16 .. 2^24 samples, H = 200 (model NL, DT = 0.8/200, FP32, generate mode),
sharded over the ranks of one box by contiguous global sample index, one exchange per control step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/sweep_mgpu.py [peer|nccl]
    python tools/sweep_mgpu.py            # one GPU

Prints one JSON line per K on rank 0: device-resident closed loop, CUDA events around 20 steps, max over ranks."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from mpc_rs_b200 import Mppi, models
    from mpc_rs_b200 import _abi as A
    from mpc_rs_b200 import distributed as D
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    transport = sys.argv[1] if len(sys.argv) > 1 else "peer"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    H, steps = 200, 20
    for lg in (16, 18, 20, 22, 24):
        K = 1 << lg
        m = Mppi(H, K, model=models.NL, lam=0.5, std_dev=3.0, limit=(-20.0, 20.0), precision="f32", dt=0.8 / H, device=local, rank=rank,
                 world_size=world, seed=20240005)
        if world > 1:
            (D.attach_mppi_peers if transport == "peer" else D.attach_mppi)(m)
        d = [C.c_void_p() for _ in range(3)]
        for q, n in zip(d, (32, 8 * H, 8 * H)):
            A.check(A.lib().mpcb_device_alloc(local, n, C.byref(q)))
        x0, u0 = np.array([0.5, 0.0, 0.1, 0.0]), np.zeros(H)
        A.check(A.lib().mpcb_device_upload(local, d[0], x0.ctypes.data_as(C.c_void_p), 32))
        A.check(A.lib().mpcb_device_upload(local, d[1], u0.ctypes.data_as(C.c_void_p), 8 * H))
        stream = torch.cuda.ExternalStream(m.stream, device=local)
        for i in range(5):
            m.compute_device(d[0].value, d[1 + (i & 1)].value, d[1 + ((i + 1) & 1)].value)
        m.sync()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            m.compute_device(d[0].value, d[1 + ((i + 1) & 1)].value, d[1 + (i & 1)].value)
        e1.record(stream)
        m.sync()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        info = m.last_info()[0]
        if rank == 0:
            print(json.dumps({"K": K, "H": H, "n_gpus": world, "transport": transport if world > 1 else "none", "ms_per_step": ms,
                              "rollout_steps_per_sec": K * H / (ms * 1e-3), "status": info["status"], "n_finite": info["n_finite"]}), flush=True)
        for q in d:
            A.lib().mpcb_device_free(local, q)
        m.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
