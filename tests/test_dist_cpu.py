"""N > 1 host logic on CPU: world_size-2 (and 3) gloo processes exercise the partition arithmetic, the
unique-id hand-off, the replica assignment and the halo message plan (with the oracle standing in for
the slabs: two half-domain oracles exchanging one column per step must reproduce the undivided oracle)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from incerto_b200 import dist as D
    from incerto_b200 import partition_range        # the engine's own arithmetic (no GPU needed)
    W = 1000
    mine = partition_range(W, rank, world)
    assert mine == D.partition_ranges(W, world)[rank]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    assert gathered[0][0] == 0 and gathered[-1][1] == W and all(gathered[i][1] == gathered[i + 1][0] for i in range(world - 1))
    uid = D.share_unique_id(dist, lambda: bytes(range(128)), rank)
    assert uid == bytes(range(128))
    reps = D.replica_assignment(10, rank, world)
    allr = [None] * world
    dist.all_gather_object(allr, reps)
    assert sorted(sum(allr, [])) == list(range(10))
    tot = D.reduce_replica_counts(dist, np.array([rank + 1, 10 * (rank + 1)]), torch)
    assert tot.tolist() == [world * (world + 1) // 2, 10 * world * (world + 1) // 2]
    # halo message plan: H cell bytes + overlay trailer, to the lower and upper neighbour
    H, n_ov = 64, 3
    msg = np.full(H + n_ov, rank, np.uint8)
    recv_lo, recv_hi = np.zeros_like(msg), np.zeros_like(msg)
    ops = []
    if rank > 0:
        ops += [dist.P2POp(dist.isend, torch.from_numpy(msg), rank - 1), dist.P2POp(dist.irecv, torch.from_numpy(recv_lo), rank - 1)]
    if rank + 1 < world:
        ops += [dist.P2POp(dist.isend, torch.from_numpy(msg), rank + 1), dist.P2POp(dist.irecv, torch.from_numpy(recv_hi), rank + 1)]
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    if rank > 0:
        assert (recv_lo == rank - 1).all()
    if rank + 1 < world:
        assert (recv_hi == rank + 1).all()
    dist.barrier()
    dist.destroy_process_group()
    q.put(rank)


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_host_logic(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(world)) == list(range(world))
