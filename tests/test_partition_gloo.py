"""The N>1 path on CPU: world_size 2 over gloo.  Every rank segments its shard of a batch (the oracle
stands in for the GPU here -- this tests the partition / gather logic, not the kernels) and the
gathered result must equal the single-process result image by image."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from mergenet_b200 import partition


def test_shard_range_matches_array_split():
    for n in (0, 1, 5, 8, 17, 256):
        for world in (1, 2, 3, 4, 8):
            ref = np.array_split(np.arange(n), world)
            for r in range(world):
                a, b = partition.shard_range(n, world, r)
                assert list(range(a, b)) == list(ref[r]), (n, world, r)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    a, b = partition.shard_range(n_items, world, rank)
    counts = []
    for i in range(a, b):
        cp, sp, C, offs = cases.cityscapes_like(24, 32, 100 + i, i % 2 == 0)
        m, oc, _ = oracle.oracle_run_segmentation(cp, sp, C, offs, *cases.RECIPE_OPTS)
        counts.append(len(oc))
    got = partition.gather_counts(torch.tensor(counts, dtype=torch.int32), n_items, world, rank)
    if rank == 0:
        q.put(got.tolist())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_partition_and_gather_equals_single_process():
    import oracle
    n_items, world = 5, 2
    expect = []
    for i in range(n_items):
        cp, sp, C, offs = cases.cityscapes_like(24, 32, 100 + i, i % 2 == 0)
        m, oc, _ = oracle.oracle_run_segmentation(cp, sp, C, offs, *cases.RECIPE_OPTS)
        expect.append(len(oc))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got == expect
