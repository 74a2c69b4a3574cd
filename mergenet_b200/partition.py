"""Batch partitioning across GPUs (north star (4)): images are independent, so a batch is split by
image index exactly like the reference shards a dataset across jobs
(``utils/dataset.py:56-63``: ``np.array_split(ids, num_jobs)[job - 1]``) and the hot path needs no
collective.  The only exchange is the final result gather, done with ``torch.distributed`` (NCCL on
GPUs, gloo in the CPU test-suite)."""
import numpy as np


def shard_range(n_items, world_size, rank):
    """[start, stop) of ``np.array_split(range(n_items), world_size)[rank]``."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    base, extra = divmod(int(n_items), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_indices(n_items, world_size, rank):
    a, b = shard_range(n_items, world_size, rank)
    return np.arange(a, b)


def gather_counts(local_counts, n_items, world_size, rank, device=None):
    """All-gather the per-image instance counts of every rank's shard into one int32[n_items] array
    (the "score/result gather").  ``local_counts``: int32 tensor of this rank's shard."""
    import torch
    import torch.distributed as dist
    sizes = [shard_range(n_items, world_size, r) for r in range(world_size)]
    longest = max(b - a for a, b in sizes)
    dev = local_counts.device if device is None else device
    pad = torch.full((longest,), -1, dtype=torch.int32, device=dev)
    pad[:local_counts.numel()] = local_counts.to(torch.int32)
    bufs = [torch.empty_like(pad) for _ in range(world_size)]
    dist.all_gather(bufs, pad)
    out = torch.empty((n_items,), dtype=torch.int32, device=dev)
    for r, (a, b) in enumerate(sizes):
        out[a:b] = bufs[r][:b - a]
    return out
