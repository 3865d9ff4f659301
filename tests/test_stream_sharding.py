"""CPU, world_size 2 over gloo: the N>1 host logic of the frame-parallel benchmark (sharding + max-over-ranks timing)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stereo_match_traditional_b200.stream import aggregate_fps, frame_seed, frames_for_rank, reduce_max


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = [f for s in range(3) for f in frames_for_rank(rank, world, 2, s)]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    slow = reduce_max(10.0 + 5.0 * rank)          # rank 1 is the slow one
    dist.barrier()
    q.put((rank, mine, gathered, slow))
    dist.destroy_process_group()


def test_two_rank_sharding_and_max_reduction():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, mine, gathered, slow in res:
        allf = sorted(f for g in gathered for f in g)
        assert allf == list(range(3 * world * 2)), "frames must be disjoint and cover the stream"
        assert slow == 15.0, "time is the max over ranks"
    assert aggregate_fps(2, 2, 3, 15.0) == pytest.approx(2 * 2 * 3 / 0.015)
    assert frame_seed(0) == 1004


def test_single_process_identity():
    assert reduce_max(3.5) == 3.5
    assert frames_for_rank(0, 1, 4, 2) == [8, 9, 10, 11]
