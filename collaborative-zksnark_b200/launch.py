"""One MPC party per rank: rendezvous and witness distribution (the Python mirror of
MpcMultiNet::init_from_file + Reveal::king_share_batch, mpc-net/src/multi.rs:51-141,
mpc-algebra/src/share/spdz.rs:150-162).

Control plane: torch.distributed (gloo, CPU tensors) carries the 128-byte NCCL unique id, the king's
scatter of witness shares (outside the timed section, like the reference) and the max-over-ranks of the
timings.  Data plane: the NCCL communicator owned by libczk_b200 (czk_net_*), one rank per GPU.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from . import binding


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_control_plane():
    """Join the gloo group described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun)."""
    rank, world, _ = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo", rank=rank, world_size=world)
    return rank, world


def broadcast_bytes_from_king(data: bytes | None, nbytes: int) -> bytes:
    """King's byte string to every party (control plane)."""
    rank, world, _ = env_rank()
    if world == 1:
        return data
    t = torch.zeros(nbytes, dtype=torch.uint8)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    return bytes(t.numpy().tobytes())


def king_share_scatter(values_mont: np.ndarray | None, k: int, seed: int) -> np.ndarray:
    """king_share_batch: the king (rank 0) splits k secrets into additive shares and sends party i its slice
    (recv_from_king, mpc-net/src/multi.rs:211-242).  Returns this party's (k, 4) share array."""
    rank, world, _ = env_rank()
    if world == 1:
        return binding.king_share_batch(values_mont, 1, seed)[0]
    mine = torch.empty((k, 4), dtype=torch.int64)
    if rank == 0:
        shares = binding.king_share_batch(values_mont, world, seed)
        chunks = [torch.from_numpy(shares[p].view(np.int64).copy()) for p in range(world)]
        dist.scatter(mine, scatter_list=chunks, src=0)
    else:
        dist.scatter(mine, src=0)
    return mine.numpy().view(np.uint64)


def max_over_ranks(x: float) -> float:
    rank, world, _ = env_rank()
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized():
        dist.barrier()


class Party:
    """Context + network for this rank: party id = rank, king = rank 0, GPU = LOCAL_RANK."""

    def __init__(self):
        self.rank, self.world = init_control_plane()
        _, _, local = env_rank()
        self.ctx = binding.Context(local)
        uid = None
        if self.world > 1:
            uid = broadcast_bytes_from_king(self.ctx.net_unique_id() if self.rank == 0 else None, 128)
        self.ctx.net_init(self.rank, self.world, uid)

    def close(self):
        self.ctx.close()
        if dist.is_initialized():
            dist.destroy_process_group()
