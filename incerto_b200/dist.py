"""Host-side multi-GPU plumbing (one process per GPU).  torch.distributed only carries the NCCL unique id
and reduces small host arrays; the data path (halo columns, aggregate all-reduce) runs inside the engine."""
from __future__ import annotations

import numpy as np


def partition_ranges(width: int, world: int) -> list[tuple[int, int]]:
    """Columns owned by each rank; same arithmetic as incerto_partition_range (csrc/module_forest.cu)."""
    return [(width * r // world, width * (r + 1) // world) for r in range(world)]


def owner_of_column(width: int, world: int, x: int) -> int:
    for r, (b, e) in enumerate(partition_ranges(width, world)):
        if b <= x < e:
            return r
    raise ValueError(x)


def replica_assignment(n_replicas: int, rank: int, world: int) -> list[int]:
    """Independent Monte Carlo replicas: replica r runs on rank r mod world."""
    return list(range(rank, n_replicas, world))


def share_unique_id(dist, make_id, rank: int) -> bytes:
    """Rank 0 creates the 128-byte NCCL unique id; everybody receives it."""
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    assert isinstance(box[0], (bytes, bytearray)) and len(box[0]) == 128
    return bytes(box[0])


def reduce_replica_counts(dist, local: np.ndarray, torch) -> np.ndarray:
    """Sum of per-replica count vectors over all ranks (host fallback used by tests; the engine's
    incerto_sim_allreduce_u64 does the same over NCCL)."""
    t = torch.from_numpy(np.ascontiguousarray(local, dtype=np.int64))
    dist.all_reduce(t)
    return t.numpy()
