import sys, time
import numpy as np
sys.path.insert(0, ".")
from mpc_rs_b200.closed_loop import ClosedLoopBatch, _upload, _download
from mpc_rs_b200 import _abi as A
for prec in ("f32", "f64"):
    loop = ClosedLoopBatch(4096, 8192, precision=prec)
    for _ in range(3): loop.tick()
    T = {}
    def tm(name, f):
        t0 = time.perf_counter(); r = f(); T[name] = T.get(name, 0) + time.perf_counter() - t0; return r
    n = 10
    for _ in range(n):
        tm("plant", lambda: setattr(loop, "x", loop.plant.dynamics_short(loop.x, loop.u0, loop.tick_dt, 0.0)))
        z = tm("sensor", lambda: loop.plant.sensor(loop.x, loop.rng))
        tm("upload z", lambda: _upload(loop.dev, loop.d_z, z.T))
        tm("ukf launch+gather+sync", lambda: (loop.ukf.run_device(1, loop.d_z, d_u=loop.d_u0, dt=loop.tick_dt), loop.ukf.gather_state_device((0, 1, 3, 4), loop.d_x4), loop.ukf.sync()))
        nxt = loop.cur ^ 1
        tm("mppi launch+sync", lambda: (loop.mppi.compute_device(loop.d_x4, loop.d_u[loop.cur], loop.d_u[nxt]), loop.mppi.first_control_device(loop.d_u[nxt], loop.d_u0), loop.mppi.sync()))
        loop.cur = nxt
        tm("download u0", lambda: _download(loop.dev, loop.d_u0, loop.u0))
    print(prec, {k: f"{v / n * 1e3:.3f} ms" for k, v in T.items()}, "block", loop.mppi.cfg.precision)
    loop.close()
