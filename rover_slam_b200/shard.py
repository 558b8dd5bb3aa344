"""Data-parallel sharding of an independent frame-pair stream over ranks.

The path has no cross-unit reduction: rank r owns a contiguous block of pairs (both frames of a pair stay on one
GPU so descriptors never cross NVLink); the only collectives are the ingest scatter of the u8 frames from rank 0 and
the gather of per-pair match counts.  Works with the nccl backend on GPUs and with gloo on CPU (tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def pair_range(n_pairs: int, rank: int, world: int):
    """Contiguous block of ceil(n_pairs / world) pairs per rank (the last ranks may own fewer / none)."""
    per = (n_pairs + world - 1) // world
    lo = min(rank * per, n_pairs)
    return lo, min(lo + per, n_pairs)


def scatter_pairs(frames, n_pairs: int, shape, device, src: int = 0):
    """frames: uint8 tensor [n_pairs, 2, H, W] on rank `src` (ignored elsewhere).  Returns this rank's
    [per, 2, H, W] block on `device`, zero-padded to ceil(n_pairs / world) pairs, and the number of valid pairs."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    per = (n_pairs + world - 1) // world
    lo, hi = pair_range(n_pairs, rank, world)
    out = torch.zeros((per, 2) + tuple(shape), dtype=torch.uint8, device=device)
    if world == 1:
        out[: hi - lo].copy_(frames[lo:hi])
        return out, hi - lo
    chunks = None
    if rank == src:
        chunks = []
        for r in range(world):
            a, b = pair_range(n_pairs, r, world)
            c = torch.zeros_like(out)
            c[: b - a].copy_(frames[a:b])
            chunks.append(c)
    dist.scatter(out, chunks, src=src)
    return out, hi - lo


def gather_counts(local_counts: torch.Tensor, n_pairs: int):
    """All ranks contribute their per-pair counts (padded to `per`); returns the [n_pairs] vector on every rank."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local_counts[:n_pairs].clone()
    parts = [torch.zeros_like(local_counts) for _ in range(world)]
    dist.all_gather(parts, local_counts)
    return torch.cat(parts)[:n_pairs]
