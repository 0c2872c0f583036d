"""ONE air_traffic_3d world split by x over several ranks, with aircraft migrating between the slabs
and ghost aircraft across the cuts - here on one GPU: the slabs of all ranks live side by side in one process and the
per-step mailboxes travel through the manual transport (incerto_sim_halo_export / _import), so the ownership, ghost and
migration logic of the kernels is checked against the oracle without several devices (the NCCL transport is checked by
bench.py --gpus N: `air_partition_parity`)."""
import numpy as np
import pytest

import incerto_b200 as ib
from incerto_b200 import _ffi as F
from conftest import SEED
from oracle.air import AirOracle

pytestmark = pytest.mark.gpu


def build_rank(n, size, height, rank, world, replica=0, host=None, systems=3):
    b = ib.SimulationBuilder(seed=SEED, replica=replica)
    position = b.component("GridPosition3D", F.I16, lanes=3)
    target = b.component("Aircraft.target_altitude", F.I16)
    b.set_row_partition(rank, world)
    b.add_spatial_grid_3d(position, target, bounds=((0, 0, 0), (size, size, height)))
    if host is None:
        b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_AIRCRAFT, F.SpawnAircraftParams(position.handle, target.handle, n, size, height)))
    else:
        xyz, tgt = host
        b.add_entity_spawner(lambda s: s.spawn_batch(len(tgt), {position: xyz, target: tgt}))
    sysl = []
    if systems & 2:
        sysl.append(ib.MoveAircraft(position, target, size, height))
    if systems & 1:
        sysl.append(ib.CheckConflicts(position, target, size, height))
    b.add_systems(*sysl)
    return b.build(), position, target


def step_all(sims):
    for s, _, _ in sims:
        s.run(1)
    # rank r's upper outbox -> rank r+1's lower inbox, rank r+1's lower outbox -> rank r's upper inbox
    for r in range(len(sims) - 1):
        lo, hi = sims[r][0], sims[r + 1][0]
        hi.halo_import(0, lo.halo_export(1))
        lo.halo_import(1, hi.halo_export(0))


def gather(sims, n, conflicts=True):
    """positions of all aircraft by uid (every aircraft must be owned by exactly one rank), conflict ends summed"""
    xyz = np.full((n, 3), -999, np.int32)
    owner = np.full(n, -1, np.int32)
    ends = 0
    for r, (s, position, target) in enumerate(sims):
        p, uids = s.read_component(position, with_uids=True)
        assert (owner[uids] == -1).all(), "an aircraft is owned by two ranks"
        owner[uids] = r
        xyz[uids] = p.astype(np.int32)
        if len(uids) and conflicts:
            ends += int(s.sample_aggregate(ib.AirConflicts(target)))
    assert (owner >= 0).all(), "an aircraft is owned by no rank"
    return xyz, owner, ends


@pytest.mark.parametrize("world,n,size,height,replica", [(2, 4000, 20, 10, 0), (3, 9000, 30, 6, 1), (4, 30_000, 64, 10, 2), (8, 20_000, 16, 4, 3)])
def test_partitioned_airspace_matches_oracle(world, n, size, height, replica):
    sims = [build_rank(n, size, height, r, world, replica=replica) for r in range(world)]
    o = AirOracle(SEED, replica, n, size, height).spawn()
    xyz0, owner0, _ = gather(sims, n, conflicts=False)     # (no step has run yet: nothing to count)
    assert np.array_equal(xyz0, o.state()[0])
    moved_ranks = 0
    for k in range(40):
        step_all(sims)
        o.run(1)
        xyz, owner, ends = gather(sims, n)
        assert np.array_equal(xyz, o.state()[0]), f"step {k + 1}"
        assert ends == 2 * int(o.series()[-1]), f"step {k + 1}: conflicts"
        # ownership follows the position
        ext = size + 1
        for r in range(world):
            b, e = ib.partition_range(ext, r, world)
            assert ((xyz[owner == r, 0] >= b) & (xyz[owner == r, 0] < e)).all()
        moved_ranks += int((owner != owner0).sum())
        owner0 = owner
    assert moved_ranks > 0, "no aircraft ever crossed a cut: the case does not test migration"
    assert o.series().max() > 0


def test_partitioned_airspace_known_answers():
    """RNG-free (check_conflicts only): pairs straddling the cut, on the cut's two sides, and in the corners."""
    size, height, world = 9, 4, 2          # ext_x = 10: rank 0 owns x in 0..4, rank 1 owns 5..9
    xyz = np.array([[4, 3, 1], [5, 3, 1],     # face neighbours across the cut: 1 conflict
                    [4, 7, 2], [4, 7, 3],     # +-z inside rank 0's boundary column: 1
                    [5, 0, 0], [5, 1, 0],     # +-y inside rank 1's boundary column: 1
                    [4, 5, 0], [5, 6, 0],     # diagonal across the cut: none
                    [0, 0, 0], [9, 9, 4],     # far corners: none
                    [3, 3, 1]], np.int16)     # -x neighbour of the first: 1
    tgt = xyz[:, 2].copy()
    sims = [build_rank(0, size, height, r, world, host=(xyz, tgt), systems=1) for r in range(world)]
    o = AirOracle(SEED, 0, 0, size, height, systems=1).spawn_explicit(xyz, tgt)
    for _ in range(3):
        step_all(sims)
        o.run(1)
        _, owner, ends = gather(sims, len(tgt))
        assert ends == 2 * int(o.series()[-1]) == 8
        assert list(owner) == [0, 1, 0, 0, 1, 1, 0, 1, 0, 1, 0]


def test_partition_needs_two_columns_per_rank():
    with pytest.raises(ib.EngineError) as e:
        build_rank(10, 2, 3, 0, 2)
    assert e.value.code == F.ERR_INVALID_ARGUMENT
