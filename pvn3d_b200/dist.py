"""Multi-GPU plumbing: frames are independent, so the path shards by frame with no data-path
collective.  One process per GPU (torchrun); rank r owns frames r, r+W, ...;
the only communication is one all_gather of the small result tensors at the end of a batch.
Backend: NCCL on GPUs, gloo in the CPU unit tests.
"""
from __future__ import annotations

import os
from typing import List, Tuple

import torch
import torch.distributed as dist


def env_rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment (no-op for a single process)."""
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_frames(n_frames: int, rank: int, world: int) -> List[int]:
    """Static round-robin frame shard: rank r takes frames r, r+W, ...."""
    return list(range(rank, n_frames, world))


def gather_frame_results(local: torch.Tensor, n_frames: int, rank: int, world: int) -> torch.Tensor:
    """local: [n_local, ...] results of frames shard_frames(n_frames, rank, world), same trailing shape
    on every rank.  Returns [n_frames, ...] in global frame order on every rank (one all_gather of a
    padded block; the payload is ~2 kB per frame)."""
    if world == 1:
        return local
    per = (n_frames + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.size(0)] = local
    blocks = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(blocks, pad)
    out = torch.empty((n_frames,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        ids = shard_frames(n_frames, r, world)
        out[ids] = blocks[r][: len(ids)]
    return out


def max_over_ranks(value: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
