"""CPU, world_size 2 over gloo: the stream sharding rule and the final event-count all-gather -- the only
collective of the path.  On the GPU box the same code runs over NCCL."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pydvs_b200 import sharding


def test_shard_rule_covers_every_stream_once():
    for total, world in ((256, 8), (64, 8), (10, 4), (3, 4), (4096, 2), (1, 1)):
        seen = []
        for r in range(world):
            first, n = sharding.shard_streams(total, world, r)
            seen += list(range(first, first + n))
            for s in range(first, first + n):
                assert sharding.owner_of_stream(s, total, world) == r
        assert seen == list(range(total))
    assert sharding.shard_streams(256, 8, 3) == (96, 32)   # cfg 5: 32 streams per GPU at G = 8


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, n = sharding.shard_streams(total, world, rank)
        local = torch.arange(first, first + n, dtype=torch.int64) * 10 + 1   # fake per-stream event counts
        allc = sharding.gather_event_counts(local, total_streams=total)
        assert allc.tolist() == [s * 10 + 1 for s in range(total)]
        off = sharding.global_event_offsets(allc)
        assert off[0] == 0 and off[-1] == allc.sum() and off.numel() == total + 1
    finally:
        dist.destroy_process_group()


def test_event_count_allgather_world2_gloo():
    mp.spawn(_worker, args=(2, _free_port(), 7), nprocs=2, join=True)   # uneven split: 4 + 3 streams


def test_gather_without_process_group_is_identity():
    c = torch.tensor([3, 1, 4], dtype=torch.int64)
    assert sharding.gather_event_counts(c).tolist() == [3, 1, 4]
