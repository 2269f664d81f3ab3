"""Developer probe (torchrun, 2+ GPUs): where the multi-GPU step time beyond the kernel's own timeline goes.
Back-to-back device-resident steps of (a) the sharded handle with the fused peer exchange, (b) the same shard without
an exchange (compute_partial), (c) an unsharded handle created in the same process after the peers were attached."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch
import torch.distributed as dist
from mpc_rs_b200 import Mppi, models
from mpc_rs_b200 import _abi as A
from mpc_rs_b200 import distributed as D

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H, KL = 100, 65536
L = A.lib()
kw = dict(model=models.NL, lam=0.5, std_dev=3.0, limit=(-20, 20), precision="f32", dt=0.8 / H, device=local)
p = C.c_void_p()
A.check(L.mpcb_device_alloc(local, 8 * (4 + 2 * H) + 8 * 256, C.byref(p)))
xu = np.concatenate([[0.5, 0, 0.1, 0.0], np.zeros(2 * H)])
L.mpcb_device_upload(local, p, xu.ctypes.data_as(C.c_void_p), xu.nbytes)
d_x, d_u, d_o = p.value, p.value + 32, p.value + 32 + 8 * H


def b2b(fn, sync, reps=300):
    best = 1e9
    for _ in range(3):
        for _ in range(10):
            fn()
        sync()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        sync()
        best = min(best, (time.perf_counter() - t0) / reps)
    return best * 1e6


single0 = Mppi(H, KL, **kw)
t_single_before = b2b(lambda: single0.compute_device(d_x, d_u, d_o), single0.sync)
m = Mppi(H, KL * world, rank=rank, world_size=world, **kw)
D.attach_mppi_peers(m)
t_peer = b2b(lambda: m.compute_device(d_x, d_u, d_o), m.sync)
x = np.array([[0.5, 0, 0.1, 0.0]])
u = np.zeros((1, H))
d_part = C.c_void_p()
A.check(L.mpcb_device_alloc(local, 8 * m.partial_len * 4, C.byref(d_part)))
t_partial = b2b(lambda: m.compute_partial(x, u, d_part.value), m.sync, reps=100)
single = Mppi(H, KL, **kw)
t_single_after = b2b(lambda: single.compute_device(d_x, d_u, d_o), single.sync)
out = torch.tensor([t_single_before, t_peer, t_partial, t_single_after], device="cuda")
dist.all_reduce(out, op=dist.ReduceOp.MAX)
if rank == 0:
    a, b, c, d = out.tolist()
    print(f"world={world}: unsharded handle before any peer mapping {a:.1f} us | sharded + peer exchange {b:.1f} us | "
          f"sharded, no exchange (compute_partial: host-staged inputs) {c:.1f} us | unsharded handle after attach {d:.1f} us")
dist.barrier()
dist.destroy_process_group()
