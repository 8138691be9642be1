"""Host-side N>1 logic on CPU: frame partitioning and the size gather over gloo (world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cairo_zstd_b200.sharding import gather_shard_summary, partition_frames


def test_partition_is_a_balanced_exact_cover():
    rng = np.random.default_rng(3)
    costs = np.exp(rng.uniform(np.log(1024), np.log(4 << 20), size=500)).astype(np.int64)
    for w in (1, 2, 4, 8):
        shards = partition_frames(costs, w)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(costs)))
        loads = [int(costs[s].sum()) for s in shards]
        assert max(loads) - min(loads) <= int(costs.max())
    assert partition_frames([], 4) == [[], [], [], []]
    assert partition_frames([5], 2) == [[0], []]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, costs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shards = partition_frames(costs, world)
    mine = shards[rank]
    bytes_out = sum(int(costs[i]) for i in mine)
    summary = gather_shard_summary(bytes_out, bytes_out // 3, rank, dist)
    q.put((rank, mine, summary))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_over_gloo():
    costs = [int(c) for c in np.random.default_rng(1).integers(1000, 100000, size=64)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, costs, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    (r0, s0, sum0), (r1, s1, sum1) = got
    assert sorted(s0 + s1) == list(range(64)) and not set(s0) & set(s1)
    assert sum0 == sum1  # every rank sees the same gathered table
    assert sum0[0][0] + sum0[1][0] == sum(costs)
    assert [t[2] for t in sum0] == [0, 1]
