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


class PairStream:
    """BASELINE config 5 as a pipeline: the frame stream lives in (pinned) host memory on rank 0; every step rank 0 copies one
    step's frames of ALL ranks to its device, scatters one block per rank, every rank runs `compute` on its block and the
    fixed-size result records are gathered back to rank 0 and copied to the host.  Ingest of step i+1 and the result gather
    of step i-1 run on a side stream while step i computes (two buffer sets, CUDA events); with the gloo backend on CPU
    the same code runs without streams (tests/test_shard_gloo.py).

      block_shape : shape of one rank's input block per step, e.g. (2*P, H, W) uint8
      result_words: int32 words of one rank's packed result record per step
      compute(in_block, out_record, step): enqueue the work of one step on the CURRENT stream
    """

    def __init__(self, block_shape, result_words: int, device, compute, src: int = 0):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.dev, self.src, self.compute = torch.device(device), src, compute
        self.cuda = self.dev.type == "cuda"
        self.block_shape, self.words = tuple(block_shape), int(result_words)
        mk = lambda shape, dt: [torch.zeros(shape, dtype=dt, device=self.dev) for _ in range(2)]
        self.in_buf = mk(self.block_shape, torch.uint8)
        self.out_buf = mk((self.words,), torch.int32)
        root = self.rank == src
        self.stage_in = mk((self.world,) + self.block_shape, torch.uint8) if root else None
        self.gathered = mk((self.world, self.words), torch.int32) if root else None
        self.host_out = None
        if root:
            self.host_out = [torch.zeros((self.world, self.words), dtype=torch.int32, pin_memory=self.cuda) for _ in range(2)]
        self.h2d_bytes = self.d2h_bytes = self.collective_bytes = 0
        if self.cuda:
            self.main = torch.cuda.current_stream(self.dev)
            self.comm = torch.cuda.Stream(self.dev)
            ev = lambda: [torch.cuda.Event() for _ in range(2)]
            self.ev_in, self.ev_out, self.ev_free_in, self.ev_free_out, self.ev_done = ev(), ev(), ev(), ev(), ev()
            for e in self.ev_free_in + self.ev_free_out:
                e.record(self.main)

    class _Null:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    def _on_comm(self):
        return torch.cuda.stream(self.comm) if self.cuda else PairStream._Null()

    def _ingest(self, i, host_step):
        """host_step: [world, *block_shape] uint8 host tensor on the source rank (None elsewhere)."""
        j = i & 1
        with self._on_comm():
            if self.cuda:
                self.comm.wait_event(self.ev_free_in[j])        # step i-2 has finished reading in_buf[j]
            chunks = None
            if self.rank == self.src:
                self.stage_in[j].copy_(host_step, non_blocking=True)
                self.h2d_bytes += host_step.numel()
                chunks = list(self.stage_in[j].unbind(0))
            if self.world > 1:
                dist.scatter(self.in_buf[j], chunks, src=self.src)
                if self.rank == self.src:
                    self.collective_bytes += host_step.numel()
            else:
                self.in_buf[j].copy_(chunks[0], non_blocking=True)
            if self.cuda:
                self.ev_in[j].record(self.comm)

    def _compute(self, i):
        j = i & 1
        if self.cuda:
            self.main.wait_event(self.ev_in[j])
            self.main.wait_event(self.ev_free_out[j])           # the gather of step i-2 has read out_buf[j]
        self.compute(self.in_buf[j], self.out_buf[j], i)
        if self.cuda:
            self.ev_free_in[j].record(self.main)
            self.ev_out[j].record(self.main)

    def _egress(self, i):
        j = i & 1
        with self._on_comm():
            if self.cuda:
                self.comm.wait_event(self.ev_out[j])
            if self.world > 1:
                dist.gather(self.out_buf[j], list(self.gathered[j].unbind(0)) if self.rank == self.src else None, dst=self.src)
            elif self.rank == self.src:
                self.gathered[j][0].copy_(self.out_buf[j], non_blocking=True)
            if self.rank == self.src:
                self.host_out[j].copy_(self.gathered[j], non_blocking=True)
                self.d2h_bytes += self.gathered[j].numel() * 4
                self.collective_bytes += self.gathered[j].numel() * 4 if self.world > 1 else 0
            if self.cuda:
                self.ev_free_out[j].record(self.comm)
                self.ev_done[j].record(self.comm)

    def run(self, n_steps: int, host_steps, consume=None):
        """host_steps(i) -> the step's [world, *block_shape] host tensor on the source rank.  consume(i, host_record) is called
        on the source rank once step i's records have landed in host memory ([world, result_words] int32)."""
        get = (lambda i: host_steps(i)) if self.rank == self.src else (lambda i: None)
        self._ingest(0, get(0))
        for i in range(n_steps):
            if i + 1 < n_steps:
                self._ingest(i + 1, get(i + 1))
            self._compute(i)
            self._egress(i)
            if i >= 1:
                self._finish(i - 1, consume)
        if n_steps:
            self._finish(n_steps - 1, consume)

    def _finish(self, i, consume):
        j = i & 1
        if self.cuda:
            self.ev_done[j].synchronize()
        if consume is not None and self.rank == self.src:
            consume(i, self.host_out[j])
