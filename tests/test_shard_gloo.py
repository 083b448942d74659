"""N>1 host logic on CPU: world_size-2 gloo processes shard files and merge detections on rank 0."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from birda_b200.shard import gather_results, shard_files


def test_shard_files_lpt_balanced_and_deterministic():
    dur = [600.0] * 10 + [3600.0, 10.0, 1800.0]
    s = shard_files(dur, 4)
    assert sorted(i for part in s for i in part) == list(range(len(dur)))
    loads = [sum(dur[i] for i in part) for part in s]
    assert max(loads) <= 3600.0 + 1e-9 and s == shard_files(dur, 4)
    assert shard_files([], 3) == [[], [], []]
    assert shard_files([1.0, 2.0], 1) == [[1, 0]]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dur = [30.0, 10.0, 20.0, 40.0, 5.0]
    mine = shard_files(dur, world)[rank]
    # fake per-file results: (start_time, confidence) tuples already sorted the reference's way
    local = {i: sorted([(float(k), 0.5 + 0.01 * i) for k in range(3)], key=lambda d: (d[0], -d[1])) for i in mine}
    merged = gather_results(local, dist)
    dist.barrier()
    if rank == 0:
        q.put({k: v for k, v in merged.items()})
    dist.destroy_process_group()


def test_two_rank_gather_on_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(merged) == [0, 1, 2, 3, 4]
    assert all(len(v) == 3 for v in merged.values())
