"""File-level sharding across the GPUs of one box and the host-side merge of detections.

The path has no exchange step: files are independent, so ranks never talk during
processing; `torch.distributed` (any backend) is used only to gather the per-file results on rank 0.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple


def shard_files(durations: Sequence[float], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of file indices to ranks (deterministic)."""
    order = sorted(range(len(durations)), key=lambda i: (-durations[i], i))
    load = [0.0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += durations[i]
    return out


def gather_results(local: Dict[int, list], dist=None) -> Dict[int, list]:
    """All ranks' {file index: sorted detections} on rank 0 (empty dict elsewhere)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(local)
    world, rank = dist.get_world_size(), dist.get_rank()
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(local, bucket, dst=0)
    merged: Dict[int, list] = {}
    if rank == 0:
        for part in bucket:
            for k, v in part.items():
                if k in merged:
                    raise RuntimeError(f"file {k} was processed by two ranks")
                merged[k] = v
    return merged
