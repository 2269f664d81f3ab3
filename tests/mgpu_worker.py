"""Multi-GPU MPPI check, one rank per GPU (launched by torchrun from tests/test_multigpu_gpu.py or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tests/mgpu_worker.py

Every rank owns a contiguous shard of the K global samples.  Checks, on every rank:
  1. replay mode, sharded over the world  == the oracle on all K samples (f64 1e-9, f32 1e-5, argmin exact)
  2. generate mode, sharded               == generate mode on ONE GPU with the same K and seed (the Philox counter
     uses the global sample index, so the drawn sample set does not depend on the world size)
  3. every rank ends with the identical u_out (bitwise)
for each exchange transport the library has ("nccl": one ncclAllGather per step; "peer": the fused peer-write
exchange over NVLink, when the build has it).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    import oracle_lib as O
    from mpc_rs_b200 import Mppi, models
    from mpc_rs_b200 import distributed as D

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    transports = ["nccl"] + (["peer"] if hasattr(D, "attach_mppi_peers") else [])

    K, H, dt, lam, sig, lim = 8192 + 40, 24, 0.03, 0.5, 3.0, (-20.0, 20.0)  # ragged: K not a multiple of anything
    x0, u0 = np.array([0.5, 0.0, 0.1, 0.0]), np.linspace(-1.0, 1.0, H)
    rng = np.random.Generator(np.random.PCG64(20240005))
    eps = sig * rng.standard_normal((K, H))
    p = O.model_defaults(O.MODEL_NL, dt=dt)
    st, u_ref, info_ref, _ = O.mppi_compute(O.MODEL_NL, p, K, H, lam, sig, lim[0], lim[1], x0, u0, eps)
    assert st == 0

    def gather_equal(u, what):
        t = torch.from_numpy(np.ascontiguousarray(u)).cuda()
        all_u = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(all_u, t)
        for r in range(world):
            assert torch.equal(all_u[r], all_u[0]), f"{what}: rank {r} disagrees with rank 0"

    for tr in transports:
        for prec, tol in (("f64", 1e-9), ("f32", 1e-5)):
            with Mppi(H, K, model=models.NL, lam=lam, std_dev=sig, limit=lim, precision=prec, dt=dt, device=local,
                      rank=rank, world_size=world, seed=77) as m, \
                 Mppi(H, K, model=models.NL, lam=lam, std_dev=sig, limit=lim, precision=prec, dt=dt, device=local,
                      seed=77) as single:
                if tr == "nccl":
                    D.attach_mppi(m)
                else:
                    D.attach_mppi_peers(m)
                for it in range(3):  # repeated steps: the exchange buffers/flags are reused
                    u = m.compute_replay(x0, u0, eps.astype(np.float32 if prec == "f32" else np.float64))
                    err = np.linalg.norm(u - u_ref) / np.linalg.norm(u_ref)
                    assert err < tol, f"[{tr}/{prec}] sharded replay vs oracle: {err}"
                    assert m.info[0]["argmax"] == info_ref["argmax"], f"[{tr}/{prec}] argmin"
                    assert m.info[0]["n_finite"] == info_ref["n_finite"]
                    gather_equal(u, f"[{tr}/{prec}] replay")
                    # keeps the single-GPU handle's call counter (part of the Philox key) in step with the sharded one
                    single.compute_replay(x0, u0, eps.astype(np.float32 if prec == "f32" else np.float64))
                for it in range(3):
                    ug = m.compute(x0, u0)
                    us = single.compute(x0, u0)
                    err = np.linalg.norm(ug - us) / np.linalg.norm(us)
                    # same samples, different block boundaries: f64 differs by summation order only, f32 by the
                    # block-local float weighted sums
                    assert err < (1e-12 if prec == "f64" else 1e-5), f"[{tr}/{prec}] sharded vs single-GPU generate: {err}"
                    assert m.last_call_info()[0]["argmax"] == single.last_call_info()[0]["argmax"]
                    gather_equal(ug, f"[{tr}/{prec}] generate")
                # device-resident closed loop (what bench.py times)
                from mpc_rs_b200 import _abi as A
                import ctypes as C
                d = [C.c_void_p() for _ in range(3)]
                for q, n in zip(d, (32, 8 * H, 8 * H)):
                    A.check(A.lib().mpcb_device_alloc(local, n, C.byref(q)))
                A.check(A.lib().mpcb_device_upload(local, d[0], x0.ctypes.data_as(C.c_void_p), 32))
                A.check(A.lib().mpcb_device_upload(local, d[1], u0.ctypes.data_as(C.c_void_p), 8 * H))
                for it in range(4):
                    m.compute_device(d[0].value, d[1 + (it & 1)].value, d[1 + ((it + 1) & 1)].value)
                m.sync()
                assert m.last_info()[0]["status"] == 0
                out = np.empty(H)
                A.check(A.lib().mpcb_device_download(local, out.ctypes.data_as(C.c_void_p), d[1], 8 * H))
                assert np.all(np.isfinite(out))
                gather_equal(out, f"[{tr}/{prec}] device loop")
                for q in d:
                    A.lib().mpcb_device_free(local, q)
            dist.barrier()
            if rank == 0:
                print(f"mgpu ok: transport={tr} precision={prec} world={world}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
