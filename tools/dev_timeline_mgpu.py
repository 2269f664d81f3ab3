"""Developer probe: per-block timeline of one sharded MPPI launch with the fused peer exchange (torchrun, MPCB_DEBUG_TS=1)."""
import ctypes as C
import os
import sys

import numpy as np

os.environ["MPCB_DEBUG_TS"] = "1"
sys.path.insert(0, ".")
import torch
import torch.distributed as dist
from mpc_rs_b200 import Mppi, models
from mpc_rs_b200 import _abi as A
from mpc_rs_b200 import distributed as D

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
K, H = 65536 * world, 100
L = A.lib()
L.mpcb_mppi_debug_timeline.restype = C.c_int64
L.mpcb_mppi_debug_timeline.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
m = Mppi(H, K, model=models.NL, lam=0.5, std_dev=3.0, limit=(-20, 20), precision="f32", dt=0.8 / H, device=local, rank=rank, world_size=world)
D.attach_mppi_peers(m)
p = C.c_void_p()
A.check(L.mpcb_device_alloc(local, 8 * (4 + 2 * H), C.byref(p)))
xu = np.concatenate([[0.5, 0, 0.1, 0.0], np.zeros(2 * H)])
L.mpcb_device_upload(local, p, xu.ctypes.data_as(C.c_void_p), xu.nbytes)
for _ in range(6):
    m.compute_device(p.value, p.value + 32, p.value + 32 + 8 * H)
m.sync()
buf = np.zeros((4096, 16), dtype=np.uint64)
n = L.mpcb_mppi_debug_timeline(m._h, buf.ctypes.data_as(C.c_void_p), 4096)
ts = buf[:n].astype(np.int64)
t0 = ts[:, 0].min()
rel = np.where(ts > 0, ts - t0, -1)
if rank == 0:
    names = {1: "rollouts done", 2: "arrived", 12: "mergers: local merge start", 13: "mergers: rank row stored to peers", 14: "mergers: flags released",
             15: "mergers: peers' rows arrived", 5: "mergers: combined, done"}
    print(f"world={world} K/rank=65536 H={H} blocks={n}")
    for i, nm in names.items():
        col = rel[:, i][rel[:, i] >= 0]
        if len(col):
            print(f"  {nm:36s} n={len(col):4d}  min {col.min()/1e3:7.2f}  median {np.median(col)/1e3:7.2f}  max {col.max()/1e3:7.2f} us")
dist.barrier()
dist.destroy_process_group()
