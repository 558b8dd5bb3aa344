"""CPU test of the N>1 host logic with the gloo backend (world_size 2): the pair stream is scattered from rank 0,
every pair is processed exactly once by exactly one rank, and the gathered per-pair results line up."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rover_slam_b200 import shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_pairs, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    h, w = 16, 24
    frames = None
    if rank == 0:
        g = torch.Generator().manual_seed(0)
        frames = torch.randint(0, 256, (n_pairs, 2, h, w), dtype=torch.uint8, generator=g)
    block, valid = shard.scatter_pairs(frames, n_pairs, (h, w), torch.device("cpu"))
    per = block.shape[0]
    # stand-in for extract+match: a checksum per pair (the GPU path is covered by the -m gpu tests)
    counts = torch.zeros(per, dtype=torch.int64)
    counts[:valid] = block[:valid].to(torch.int64).sum(dim=(1, 2, 3))
    allc = shard.gather_counts(counts, n_pairs)
    if rank == 0:
        ret["counts"] = allc.numpy().copy()
        ret["expected"] = frames.to(torch.int64).sum(dim=(1, 2, 3)).numpy()
    ret[f"range{rank}"] = shard.pair_range(n_pairs, rank, world)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [8, 5, 1])
def test_scatter_gather_world2(n_pairs):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_pairs, ret), nprocs=world, join=True)
    assert np.array_equal(ret["counts"], ret["expected"])
    covered = []
    for r in range(world):
        lo, hi = ret[f"range{r}"]
        covered += list(range(lo, hi))
    assert covered == list(range(n_pairs))


def test_pair_range_partitions():
    for n in (0, 1, 7, 8, 4096):
        for world in (1, 2, 4, 8):
            spans = [shard.pair_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) <= (n + world - 1) // world


def _stream_worker(rank, world, port, n_steps, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, h, w = 3, 8, 16
    steps = None
    if rank == 0:
        g = torch.Generator().manual_seed(1)
        steps = torch.randint(0, 256, (n_steps, world, 2 * P, h, w), dtype=torch.uint8, generator=g)

    def compute(in_block, out_record, step):            # stand-in for extract + match: per-pair checksums + the step number
        pair_sum = in_block.view(P, 2, h, w).to(torch.int64).sum(dim=(1, 2, 3)).to(torch.int32)
        out_record[:P].copy_(pair_sum)
        out_record[P] = step
        out_record[P + 1] = rank

    ps = shard.PairStream((2 * P, h, w), P + 2, "cpu", compute)
    got = {}
    ps.run(n_steps, (lambda i: steps[i]) if rank == 0 else None, consume=lambda i, rec: got.__setitem__(i, rec.clone()))
    if rank == 0:
        ret["ok"] = True
        for i in range(n_steps):
            want = steps[i].view(world, P, 2, h, w).to(torch.int64).sum(dim=(2, 3, 4)).to(torch.int32)
            ret["ok"] = ret["ok"] and bool(torch.equal(got[i][:, :P], want)) and got[i][:, P].tolist() == [i] * world \
                and got[i][:, P + 1].tolist() == list(range(world))
        ret["bytes"] = (ps.h2d_bytes, ps.d2h_bytes, ps.collective_bytes)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_steps", [1, 2, 5])
def test_pair_stream_pipeline_world2(n_steps):
    """The config-5 pipeline (ingest from rank 0, scatter, compute, gather, egress to rank 0's host memory) delivers every
    step's records of every rank, in order, with the double-buffered sets never mixed up."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_stream_worker, args=(world, _free_port(), n_steps, ret), nprocs=world, join=True)
    assert ret["ok"]
    h2d, d2h, coll = ret["bytes"]
    assert h2d == n_steps * world * 6 * 8 * 16 and d2h == n_steps * world * 5 * 4 and coll == h2d + d2h
