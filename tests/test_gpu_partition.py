"""Row-partitioned forest (BASELINE config "16384^2 row-split with halo exchange on 8") on ONE GPU:
the slabs of all ranks live side by side in one process and the halo columns travel through the manual
transport (incerto_sim_halo_export/import), so the slab / halo / overlay-ownership logic of the kernel is
checked against the oracle and against the undivided run without needing several devices."""
import numpy as np
import pytest

import incerto_b200 as ib
from incerto_b200 import _ffi as F
from conftest import SEED
from oracle.forest import ForestOracle

pytestmark = pytest.mark.gpu
HEALTHY, EMPTY = 0x40, 0x80


def build_rank(W, H, rank, world, host=None, density=0.85, p_regrow=0.004, fires=5):
    b = ib.SimulationBuilder(seed=SEED)
    pos = b.component("GridPosition2D", F.I16, lanes=2)
    cell = b.component("ForestCell.state", F.U8)
    b.set_row_partition(rank, world)
    b.add_spatial_grid_2d(pos, cell, bounds=((0, 0), (W - 1, H - 1)))
    if host is None:
        b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_FOREST_GRID, F.SpawnForestParams(pos.handle, cell.handle, W, H, density, 3, fires)))
    else:
        xy, st = host
        b.add_entity_spawner(lambda s: s.spawn_batch(len(st), {pos: xy.astype(np.int16), cell: st}))
    b.add_systems(ib.FireSpread(pos, cell, 0.6, 3), ib.BurnProgression(pos, cell), ib.Regrowth(pos, cell, p_regrow))
    b.record_aggregate_time_series(ib.FireStatsOf(cell), 1)
    return b.build(), cell


def step_all(sims):
    for s, _ in sims:
        s.run(1)
    # rank r's highest column -> rank r+1's lower halo; rank r+1's lowest column -> rank r's upper halo
    for r in range(len(sims) - 1):
        lo, hi = sims[r][0], sims[r + 1][0]
        hi.halo_import(0, lo.halo_export(1))
        lo.halo_import(1, hi.halo_export(0))


def gather_cells(sims, W, H, n_overlay):
    """Concatenate the slabs (each rank reports its columns then the full overlay array)."""
    base, ov = [], np.zeros(n_overlay, np.uint8)
    for rank, (s, cell) in enumerate(sims):
        st = s.read_component(cell)
        b, e = ib.partition_range(W, rank, len(sims))
        base.append(st[:(e - b) * H])
    return np.concatenate(base)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_partition_device_spawn_matches_oracle(world):
    W, H, steps = 64, 48, 60
    sims = [build_rank(W, H, r, world) for r in range(world)]
    o = ForestOracle(SEED, 0, W, H, density=0.85, p_regrow=0.004, fires=5).spawn()
    for _ in range(steps):
        step_all(sims)
        o.run(1)
        want = o.cells()[0]
        assert np.array_equal(gather_cells(sims, W, H, 5), want[:W * H])
        # FireStats: every rank holds the counts of its slab + the overlays it owns; their sum is the world's
        tot = np.zeros(5, np.int64)
        for s, cell in sims:
            fs = s.sample_aggregate(ib.FireStatsOf(cell))
            tot += np.array([fs.healthy_count, fs.burning_count, fs.burned_count, fs.empty_count, fs.total_cells])
        assert tuple(tot) == tuple(int(x) for x in o.stats())
    assert o.stats()[2] > 0.2 * W * H


def test_partition_overlays_on_the_cut():
    """Overlay entities sitting on both sides of a slab boundary (and a pair on the same cell): their fire must
    cross the cut through the overlay-state trailer of the halo message."""
    W, H, world = 32, 40, 2
    rng = np.random.default_rng(3)
    xs, ys = np.meshgrid(np.arange(W), np.arange(H), indexing="ij")
    xy = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.int32)
    st = np.where(rng.random(W * H) < 0.9, HEALTHY, EMPTY).astype(np.uint8)
    cut = W // 2
    ov_xy = np.array([[cut - 1, 10], [cut, 20], [cut, 20], [cut - 1, 30], [cut, 31], [0, 0], [W - 1, H - 1]], np.int32)
    ov_st = np.array([3, 3, 2, HEALTHY, 3, 3, 3], np.uint8)
    all_xy, all_st = np.concatenate([xy, ov_xy]), np.concatenate([st, ov_st])
    sims = [build_rank(W, H, r, world, host=(all_xy, all_st)) for r in range(world)]
    o = ForestOracle(SEED, 0, W, H, p_regrow=0.004, fires=0).spawn_explicit(all_xy, all_st)
    for _ in range(50):
        step_all(sims)
        o.run(1)
        want = o.cells()[0]
        assert np.array_equal(gather_cells(sims, W, H, len(ov_st)), want[:W * H])
        # overlay states: owner rank's copy
        for k, (x, _) in enumerate(ov_xy):
            owner = 0 if x < cut else 1
            s, cell = sims[owner]
            b, e = ib.partition_range(W, owner, world)
            got = s.read_component(cell)[(e - b) * H + k]
            assert got == want[W * H + k], (k, got, want[W * H + k])


def test_partition_equals_undivided_run():
    W, H = 96, 64
    whole, cell_w = build_rank(W, H, 0, 1)
    sims = [build_rank(W, H, r, 4) for r in range(4)]
    for _ in range(40):
        step_all(sims)
    whole.run(40)
    assert np.array_equal(gather_cells(sims, W, H, 5), whole.read_component(cell_w)[:W * H])


@pytest.mark.parametrize("W,H,world", [(48, 1100, 2), (60, 1300, 3)])
def test_partition_tall_slabs_match_oracle(W, H, world):
    """Slabs whose columns span several warp spans in y (H > 512; the shape every rank of the 16384-cell-column bench
    configuration has): states and summed FireStats against the oracle after every step."""
    steps = 150
    sims = [build_rank(W, H, r, world, density=0.9, p_regrow=0.002, fires=8) for r in range(world)]
    o = ForestOracle(SEED, 0, W, H, density=0.9, p_regrow=0.002, fires=8).spawn()
    for k in range(steps):
        step_all(sims)
        o.run(1)
        assert np.array_equal(gather_cells(sims, W, H, 8), o.cells()[0][:W * H]), f"step {k + 1}"
        tot = np.zeros(5, np.int64)
        for s, cell in sims:
            fs = s.sample_aggregate(ib.FireStatsOf(cell))
            tot += np.array([fs.healthy_count, fs.burning_count, fs.burned_count, fs.empty_count, fs.total_cells])
        assert tuple(tot) == tuple(int(x) for x in o.stats())
    assert o.stats()[2] > 0.2 * W * H


def test_partition_slab_local_host_spawn():
    """A partitioned handle may be handed only the rows it needs (its slab + one halo column per interior side, then
    all overlay entities) instead of the whole world: same result as the undivided run and as the oracle."""
    W, H, world = 48, 600, 3
    rng = np.random.default_rng(11)
    xs, ys = np.meshgrid(np.arange(W), np.arange(H), indexing="ij")
    xy = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.int32)
    st = np.where(rng.random(W * H) < 0.9, HEALTHY, EMPTY).astype(np.uint8)
    ov_xy = np.array([[15, 300], [16, 511], [16, 512], [40, 10], [0, 599]], np.int32)
    ov_st = np.array([3, 3, 3, 3, 3], np.uint8)
    sims = []
    for r in range(world):
        b, e = ib.partition_range(W, r, world)
        lo, hi = max(b - 1, 0), min(e + 1, W)
        sims.append(build_rank(W, H, r, world, host=(np.concatenate([xy[lo * H:hi * H], ov_xy]),
                                                      np.concatenate([st[lo * H:hi * H], ov_st])), p_regrow=0.002))
    o = ForestOracle(SEED, 0, W, H, p_regrow=0.002, fires=0).spawn_explicit(np.concatenate([xy, ov_xy]), np.concatenate([st, ov_st]))
    for k in range(80):
        step_all(sims)
        o.run(1)
        assert np.array_equal(gather_cells(sims, W, H, len(ov_st)), o.cells()[0][:W * H]), f"step {k + 1}"
    # a slab that is too short is refused
    with pytest.raises(ib.EngineError) as err:
        build_rank(W, H, 1, world, host=(np.concatenate([xy[:5 * H], ov_xy]), np.concatenate([st[:5 * H], ov_st])))
    assert err.value.code in (F.ERR_UNSUPPORTED, F.ERR_INVALID_ARGUMENT)
