"""Developer probe: per-block timeline of one MPPI launch (MPCB_DEBUG_TS=1)."""
import ctypes as C
import os
import sys

import numpy as np

os.environ["MPCB_DEBUG_TS"] = "1"
sys.path.insert(0, ".")
from mpc_rs_b200 import Mppi, models
from mpc_rs_b200 import _abi as A

K = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
H = int(sys.argv[2]) if len(sys.argv) > 2 else 100
L = A.lib()
L.mpcb_mppi_debug_timeline.restype = C.c_int64
L.mpcb_mppi_debug_timeline.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
m = Mppi(H, K, model=models.NL, lam=0.5, std_dev=3.0, limit=(-20, 20), precision="f32", dt=0.8 / H)
p = C.c_void_p()
A.check(L.mpcb_device_alloc(0, 8 * (4 + 2 * H), C.byref(p)))
xu = np.concatenate([[0.5, 0, 0.1, 0.0], np.zeros(2 * H)])
L.mpcb_device_upload(0, p, xu.ctypes.data_as(C.c_void_p), xu.nbytes)
for _ in range(5):
    m.compute_device(p.value, p.value + 32, p.value + 32 + 8 * H)
m.sync()
if len(sys.argv) > 3 and sys.argv[3] == "cold":
    # the stamped launch after an L2 flush (what bench.py's `value` times): a 512 MB fill on the same device, then one step
    import torch
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda:0")
    flush.fill_(1)
    torch.cuda.synchronize()
    m.compute_device(p.value, p.value + 32, p.value + 32 + 8 * H)
    m.sync()
    print("(cold: after an L2 flush)")
buf = np.zeros((4096, 16), dtype=np.uint64)
n = L.mpcb_mppi_debug_timeline(m._h, buf.ctypes.data_as(C.c_void_p), 4096)
ts = buf[:n].astype(np.int64)
t0 = ts[:, 0].min()
rel = np.where(ts > 0, ts - t0, -1)
print(f"K={K} H={H} blocks={n}")
names = ["start", "rollouts done", "ticket1", "group merged | ws: weights done", "ticket2 | ws: weighted sums done", "final done"]
names += ["", "m: entry", "m: loads issued", "m: after barrier 1 (warp argmax)", "m: tid0 accumulated", "m: outputs stored"]
names += ["ws: producer warp 0 done", "ws: block max known", "ws: consumer warp 0 done", "ws: last consumer warp done"]
for i, nm in enumerate(names):
    if not nm:
        continue
    col = rel[:, i][rel[:, i] >= 0]
    if len(col):
        print(f"  {nm:38s} n={len(col):4d}  min {col.min()/1e3:7.2f} us  median {np.median(col)/1e3:7.2f} us  max {col.max()/1e3:7.2f} us")

# blocks per SM and when each SM's last block finished its rollouts
smid = ts[:, 6] - 1
if (smid >= 0).all():
    per_sm = {}
    for b in range(n):
        per_sm.setdefault(int(smid[b]), []).append(rel[b, 1] / 1e3)
    by_load = {}
    for sm, v in per_sm.items():
        by_load.setdefault(len(v), []).append(max(v))
    print(f"  SMs used: {len(per_sm)}")
    for load in sorted(by_load):
        v = by_load[load]
        print(f"  {len(v):4d} SMs hold {load} blocks: last rollout done min {min(v):6.2f} median {np.median(v):6.2f} max {max(v):6.2f} us")
