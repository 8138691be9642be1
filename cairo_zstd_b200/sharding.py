"""Host-side sharding of a frame batch over GPUs.

Frames are independent (DecoderScratch::reset clears all state, src/decoding/scratch.cairo:42-58), so
the path shards with no data-path collective: the host partitions frames across ranks, each rank runs
the same batch call on its shard, and the only cross-rank step is gathering per-rank byte counts and
statuses.  A frame is never split across GPUs.
"""
from typing import List, Sequence


def partition_frames(costs: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy largest-first partition of frame indices by cost (compressed + decompressed bytes).
    Deterministic; returns world_size lists of indices, each sorted ascending."""
    order = sorted(range(len(costs)), key=lambda i: (-int(costs[i]), i))
    loads = [0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += int(costs[i])
    return [sorted(s) for s in shards]


def gather_shard_summary(local_bytes_out: int, local_bytes_in: int, local_failed: int, dist=None):
    """All-gather (bytes_out, bytes_in, failed) of every rank.  `dist` is torch.distributed (or None for
    a single process).  Returns a list of (bytes_out, bytes_in, failed) per rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [(int(local_bytes_out), int(local_bytes_in), int(local_failed))]
    import torch
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([local_bytes_out, local_bytes_in, local_failed], dtype=torch.int64, device=dev)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [tuple(int(v) for v in o.tolist()) for o in out]
