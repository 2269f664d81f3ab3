"""Host-side plumbing for the multi-GPU path: shard arithmetic and NCCL-id rendezvous.

One process per GPU.  MPPI shards the K samples of a controller contiguously over the ranks and needs ONE small
exchange per control step, done inside libmpc_b200 either by the rollout kernel itself over peer memory
(attach_mppi_peers: every rank's final block stores its partial row into the peers' mailboxes over NVLink and
combines the rows it receives — no collective call) or with one ncclAllGather (attach_mppi); batched UKF filters /
closed-loop controllers shard without any exchange.  torch.distributed is only the rendezvous that carries the
128-byte handles between the ranks — any backend works (gloo on CPU in the tests).
"""
from __future__ import annotations


def shard_range(total: int, rank: int, world_size: int):
    """Contiguous shard [first, first+count) of `total` items — the same integer arithmetic as mpcb_mppi_create
    (k_offset = K*rank/world, K_local = K*(rank+1)/world - k_offset), so the shards tile [0, total) exactly."""
    if not (0 <= rank < world_size):
        raise ValueError("rank outside [0, world_size)")
    first = total * rank // world_size
    return first, total * (rank + 1) // world_size - first


def exchange_unique_id(make_id, group=None) -> bytes:
    """Rank 0 calls make_id() (-> 128 bytes), everyone returns the same bytes (broadcast over torch.distributed)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    buf = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        raw = make_id()
        if len(raw) != 128:
            raise ValueError("ncclUniqueId must be 128 bytes")
        buf.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(buf, 0, group=group)
    return bytes(buf.cpu().numpy().tobytes())


def attach_mppi(mppi, group=None):
    """Gives a sharded Mppi handle (rank/world_size set at construction) its NCCL communicator."""
    from .mppi import comm_unique_id
    mppi.attach_comm(exchange_unique_id(comm_unique_id, group))
    return mppi


def exchange_handles(mine: bytes, group=None):
    """All-gathers one fixed-size byte string per rank; returns the list in rank order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return [bytes(o.cpu().numpy().tobytes()) for o in out]


def attach_mppi_peers(mppi, group=None):
    """Gives a sharded Mppi handle the peer-memory mailboxes of all ranks (fused in-kernel exchange)."""
    mppi.attach_peers(exchange_handles(mppi.peer_handle(), group))
    return mppi
