"""Volume sharding across GPUs / ranks and host-side gathering of HSP lists.

The path is embarrassingly parallel over database volumes: one process per GPU,
each owning whole volumes; there is no collective on the data path.  The only cross-rank steps are
the final gather of per-subject HSP lists to rank 0 (what BlastHSPStreamWrite merges in the
reference, core/blast_engine.c:1309) and timing reductions in bench.py.
"""
from __future__ import annotations

import numpy as np


def assign_volumes(n_volumes: int, world_size: int) -> list[list[int]]:
    """Round-robin assignment volume -> rank (volume v goes to rank v % world_size)."""
    out = [[] for _ in range(world_size)]
    for v in range(n_volumes):
        out[v % world_size].append(v)
    return out


def globalize_oids(hsps: np.ndarray, volume_oid_base: int) -> np.ndarray:
    """Volume-local oids -> database-wide oids (volumes are contiguous oid ranges)."""
    out = hsps.copy()
    out["oid"] += volume_oid_base
    return out


def gather_hsps(local: np.ndarray, dist=None) -> np.ndarray | None:
    """Gather every rank's HSP records on rank 0, ordered by (oid, list order) like a single
    sequential pass over the whole database.  Works with any torch.distributed backend (gloo/nccl)
    because it moves host objects; returns None on ranks != 0."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(local, bucket, dst=0)
    if rank != 0:
        return None
    parts = [p for p in bucket if p is not None and len(p)]
    if not parts:
        return local[:0]
    allh = np.concatenate(parts)
    order = np.argsort(allh["oid"], kind="stable")     # keeps each subject's list order
    return allh[order]
