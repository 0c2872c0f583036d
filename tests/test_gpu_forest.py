"""forest_fire (examples/forest_fire.rs) on the GPU vs the CPU oracle: bit-exact cell states and FireStats."""
import numpy as np
import pytest

import incerto_b200 as ib
from incerto_b200 import _ffi as F
from conftest import SEED
from oracle.forest import ForestOracle

pytestmark = pytest.mark.gpu

HEALTHY, EMPTY, BURNED = 0x40, 0x80, 0x00


def build_device_forest(W, H, replica=0, density=0.7, p_spread=0.6, burn=3, p_regrow=0.001, fires=3, systems=7,
                        series=1, seed=SEED):
    """forest_fire.rs:147-170 with the built-in device spawner."""
    b = ib.SimulationBuilder(seed=seed, replica=replica)
    pos = b.component("GridPosition2D", F.I16, lanes=2)
    cell = b.component("ForestCell.state", F.U8)
    b.add_spatial_grid_2d(pos, cell, bounds=((0, 0), (W - 1, H - 1)))
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_FOREST_GRID, F.SpawnForestParams(pos.handle, cell.handle, W, H, density,
                                                                                   burn, fires)))
    if systems & 1:
        b.add_systems(ib.FireSpread(pos, cell, p_spread, burn))
    if systems & 2:
        b.add_systems(ib.BurnProgression(pos, cell))
    if systems & 4:
        b.add_systems(ib.Regrowth(pos, cell, p_regrow))
    if series:
        b.record_aggregate_time_series(ib.FireStatsOf(cell), series)
    return b.build(), pos, cell


def stats_tuple(fs):
    return (fs.healthy_count, fs.burning_count, fs.burned_count, fs.empty_count, fs.total_cells)


@pytest.mark.parametrize("W,H", [(50, 50), (64, 64), (37, 53), (5, 3), (1, 1), (130, 48), (48, 130), (33, 16), (4, 100)])
def test_parity_device_spawn(W, H):
    """Reference config (50x50, :33-44) and ragged shapes: every cell identical to the oracle after each phase."""
    steps = 60
    sim, pos, cell = build_device_forest(W, H)
    o = ForestOracle(SEED, 0, W, H).spawn()
    st0, xy0 = o.cells()
    assert np.array_equal(sim.read_component(cell), st0)          # spawn parity (:239-278)
    assert np.array_equal(sim.read_component(pos), xy0.astype(np.int16))
    assert stats_tuple(sim.sample_aggregate(ib.FireStatsOf(cell))) == tuple(int(x) for x in o.stats())
    for chunk in (1, 1, 7, steps - 9):
        sim.run(chunk)
        o.run(chunk)
        assert np.array_equal(sim.read_component(cell), o.cells()[0])
        assert stats_tuple(sim.sample_aggregate(ib.FireStatsOf(cell))) == tuple(int(x) for x in o.stats())
    ts = sim.get_aggregate_time_series(ib.FireStatsOf(cell))
    assert ts.time() == list(range(1, steps + 1))
    got = np.array([stats_tuple(v) for v in ts.values()], dtype=np.uint64)
    assert np.array_equal(got, o.series())
    assert int(got[0, 4]) == W * H + 3                             # total_cells counts the 3 extra igniters (:260-278)


@pytest.mark.parametrize("replica", [1, 2, 5])
def test_parity_fire_actually_spreads(replica):
    """A denser forest so that the front crosses chunk, column-quad and CTA-tile borders."""
    W, H, steps = 96, 160, 120
    sim, pos, cell = build_device_forest(W, H, replica=replica, density=0.9, p_regrow=0.003, fires=5)
    o = ForestOracle(SEED, replica, W, H, density=0.9, p_regrow=0.003, fires=5).spawn()
    sim.run(steps)
    o.run(steps)
    st = sim.read_component(cell)
    assert np.array_equal(st, o.cells()[0])
    s = o.stats()
    assert s[2] > 0.3 * W * H                                      # the fire really burned a large area
    assert np.array_equal(np.array([stats_tuple(v) for v in sim.get_aggregate_time_series(ib.FireStatsOf(cell)).values()],
                                   dtype=np.uint64), o.series())


@pytest.mark.parametrize("systems", [1, 2, 3, 4, 5, 6])
def test_parity_system_subsets(systems):
    """Any subset of the three systems may be added (simulation_builder.rs:66-70)."""
    W, H = 40, 48
    sim, pos, cell = build_device_forest(W, H, density=0.85, p_regrow=0.02, systems=systems, series=0)
    o = ForestOracle(SEED, 0, W, H, density=0.85, p_regrow=0.02, systems=systems).spawn()
    sim.run(25)
    o.run(25)
    assert np.array_equal(sim.read_component(cell), o.cells()[0])


def test_regrowth_one_level_draw():
    """p_regrow above 1/256 uses the single 32-bit draw of the contract."""
    W, H = 32, 48
    sim, pos, cell = build_device_forest(W, H, density=0.5, p_regrow=0.05)
    o = ForestOracle(SEED, 0, W, H, density=0.5, p_regrow=0.05).spawn()
    sim.run(40)
    o.run(40)
    assert np.array_equal(sim.read_component(cell), o.cells()[0])


def test_parity_host_spawn_with_overlays():
    """Host-side spawner exactly as forest_fire.rs:234-279 does it: cells x-major, then extra burning entities on top,
    including two on the same cell, one on the border, one Healthy and one Empty overlay."""
    W, H = 24, 40
    rng = np.random.default_rng(9)
    xs, ys = np.meshgrid(np.arange(W), np.arange(H), indexing="ij")
    xy = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.int32)
    st = np.where(rng.random(W * H) < 0.85, HEALTHY, EMPTY).astype(np.uint8)
    ov_xy = np.array([[3, 3], [3, 3], [0, 0], [W - 1, H - 1], [10, 20], [11, 20], [12, 7]], np.int32)
    ov_st = np.array([3, 2, 3, 3, HEALTHY, EMPTY, BURNED], np.uint8)
    all_xy, all_st = np.concatenate([xy, ov_xy]), np.concatenate([st, ov_st])

    b = ib.SimulationBuilder(seed=SEED)
    pos = b.component("GridPosition2D", F.I16, lanes=2)
    cell = b.component("ForestCell.state", F.U8)
    b.add_spatial_grid_2d(pos, cell, bounds=((0, 0), (W - 1, H - 1)))
    b.add_entity_spawner(lambda s: s.spawn_batch(len(all_st), {pos: all_xy.astype(np.int16), cell: all_st}))
    b.add_systems(ib.FireSpread(pos, cell, 0.6, 3), ib.BurnProgression(pos, cell), ib.Regrowth(pos, cell, 0.004))
    b.record_aggregate_time_series(ib.FireStatsOf(cell), 1)
    sim = b.build()
    o = ForestOracle(SEED, 0, W, H, p_regrow=0.004, fires=0).spawn_explicit(all_xy, all_st)
    for _ in range(50):
        sim.run(1)
        o.run(1)
        got = sim.read_component(cell)
        assert np.array_equal(got, o.cells()[0])
    assert np.array_equal(np.array([stats_tuple(v) for v in sim.get_aggregate_time_series(ib.FireStatsOf(cell)).values()],
                                   dtype=np.uint64), o.series())
    g = b.last_grid
    assert sim.grid_entities_at(g, (3, 3)) == [3 * H + 3, W * H, W * H + 1]
    assert sim.grid_neighbors_of(g, (0, 0), orthogonal=True) == sorted([1, H])


def test_many_overlay_entities():
    """1500 overlay entities (igniters and bystanders) scattered over a 280 x 4096 grid (about 20 near every CTA tile) - more than one block of threads
    stages per pass, several CTA tiles in both directions: every step against the oracle."""
    W, H, n_ov = 280, 4096, 1500
    rng = np.random.default_rng(21)
    xs, ys = np.meshgrid(np.arange(W), np.arange(H), indexing="ij")
    xy = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.int32)
    st = np.where(rng.random(W * H) < 0.8, HEALTHY, EMPTY).astype(np.uint8)
    ov_xy = np.stack([rng.integers(0, W, n_ov), rng.integers(0, H, n_ov)], 1).astype(np.int32)
    ov_st = rng.choice(np.array([3, 2, 1, HEALTHY, EMPTY, BURNED], np.uint8), n_ov)
    all_xy, all_st = np.concatenate([xy, ov_xy]), np.concatenate([st, ov_st])
    b = ib.SimulationBuilder(seed=SEED)
    pos = b.component("GridPosition2D", F.I16, lanes=2)
    cell = b.component("ForestCell.state", F.U8)
    b.add_spatial_grid_2d(pos, cell, bounds=((0, 0), (W - 1, H - 1)))
    b.add_entity_spawner(lambda s: s.spawn_batch(len(all_st), {pos: all_xy.astype(np.int16), cell: all_st}))
    b.add_systems(ib.FireSpread(pos, cell, 0.6, 3), ib.BurnProgression(pos, cell), ib.Regrowth(pos, cell, 0.002))
    b.record_aggregate_time_series(ib.FireStatsOf(cell), 1)
    sim = b.build()
    o = ForestOracle(SEED, 0, W, H, p_regrow=0.002, fires=0).spawn_explicit(all_xy, all_st)
    for _ in range(12):
        sim.run(1)
        o.run(1)
        assert np.array_equal(sim.read_component(cell), o.cells()[0])
    assert np.array_equal(np.array([stats_tuple(v) for v in sim.get_aggregate_time_series(ib.FireStatsOf(cell)).values()],
                                   dtype=np.uint64), o.series())


def test_reset_and_replicas():
    sim, pos, cell = build_device_forest(64, 64, replica=4)
    sim.run(30)
    a = sim.read_component(cell).copy()
    sim.reset(4)
    sim.run(30)
    assert np.array_equal(a, sim.read_component(cell))
    sim.reset(7)
    sim.run(30)
    o = ForestOracle(SEED, 7, 64, 64).spawn().run(30)
    assert np.array_equal(sim.read_component(cell), o.cells()[0])
    assert sim.get_aggregate_time_series(ib.FireStatsOf(cell)).len() == 30


@pytest.mark.parametrize("W,H,fires", [(30, 1100, 6), (56, 2048, 8), (24, 513, 5), (7, 1537, 8), (100, 1024, 6)])
def test_parity_tall_grids_cross_the_warp_span(W, H, fires):
    """Columns longer than one warp span (512 cells in y; blockIdx.x > 0): the fire front crosses span and CTA-tile
    borders in y, so the neighbour-across-the-span loads of forest_step_kernel decide real ignitions
    (forest_fire.rs:282-329).  States and the FireStats record after every step, the series at the end."""
    steps = 160
    sim, pos, cell = build_device_forest(W, H, density=0.9, p_regrow=0.002, fires=fires)
    o = ForestOracle(SEED, 0, W, H, density=0.9, p_regrow=0.002, fires=fires).spawn()
    assert np.array_equal(sim.read_component(cell), o.cells()[0])
    crossed = False
    for k in range(steps):
        sim.run(1)
        o.run(1)
        st = sim.read_component(cell)
        assert np.array_equal(st, o.cells()[0]), f"step {k + 1}"
        assert stats_tuple(sim.sample_aggregate(ib.FireStatsOf(cell))) == tuple(int(x) for x in o.stats())
        g = st[:W * H].reshape(W, H)
        burning = (g > 0) & (g < 0x40)
        for y in range(512, H, 512):        # something burns on both sides of a span border at the same time
            crossed = crossed or bool(burning[:, y - 1].any() and burning[:, y].any())
    assert crossed, "the fire never crossed a warp-span border: the case does not test what it is for"
    assert np.array_equal(np.array([stats_tuple(v) for v in sim.get_aggregate_time_series(ib.FireStatsOf(cell)).values()],
                                   dtype=np.uint64), o.series())


def test_parity_tall_grid_graph_replay():
    """The same tall shape through run(n) with n > 16 (captured graphs of 16 steps + remainder)."""
    W, H = 40, 1300
    sim, pos, cell = build_device_forest(W, H, density=0.9, p_regrow=0.002, fires=8)
    o = ForestOracle(SEED, 0, W, H, density=0.9, p_regrow=0.002, fires=8).spawn()
    for n in (50, 33, 120):
        sim.run(n)
        o.run(n)
        assert np.array_equal(sim.read_component(cell), o.cells()[0])
    assert np.array_equal(np.array([stats_tuple(v) for v in sim.get_aggregate_time_series(ib.FireStatsOf(cell)).values()],
                                   dtype=np.uint64), o.series())


def test_full_size_against_the_oracle():
    """BASELINE config 1 (4096 x 4096, the bench.py configuration) x 3 steps: every cell and the FireStats series
    against the oracle (about 25 s of CPU)."""
    W = H = 4096
    sim, pos, cell = build_device_forest(W, H, series=1)
    o = ForestOracle(SEED, 0, W, H).spawn()
    assert np.array_equal(sim.read_component(cell), o.cells()[0])
    sim.run(3)
    o.run(3)
    assert np.array_equal(sim.read_component(cell), o.cells()[0])
    assert np.array_equal(np.array([stats_tuple(v) for v in sim.get_aggregate_time_series(ib.FireStatsOf(cell)).values()],
                                   dtype=np.uint64), o.series())


def test_full_size_properties():
    """BASELINE config 1 (4096 x 4096): size-independent invariants over 40 steps (the oracle comparison at this size
    is test_full_size_against_the_oracle)."""
    W = H = 4096
    sim, pos, cell = build_device_forest(W, H, series=1)
    fs0 = stats_tuple(sim.sample_aggregate(ib.FireStatsOf(cell)))
    assert fs0[4] == W * H + 3 and abs(fs0[0] / (W * H) - 0.7) < 1e-3
    sim.run(40)
    ts = sim.get_aggregate_time_series(ib.FireStatsOf(cell))
    vals = np.array([stats_tuple(v) for v in ts.values()], dtype=np.int64)
    assert (vals[:, :4].sum(1) == W * H + 3).all()                 # every entity is in exactly one state
    assert (np.diff(vals[:, 2]) >= 0).all()                        # Burned never reverts (no system leaves Burned)
    assert vals[-1, 1] > 0                                         # still burning
    # regrowth rate: Empty -> Healthy with p = 0.001 per step
    regrown = fs0[3] - vals[-1, 3]
    assert abs(regrown / (fs0[3] * 40 * 0.001) - 1.0) < 0.05
    st = sim.read_component(cell)
    assert stats_tuple(sim.sample_aggregate(ib.FireStatsOf(cell))) == tuple(vals[-1])
    assert int((st == HEALTHY).sum()) == vals[-1, 0] and int((st == EMPTY).sum()) == vals[-1, 3]


def test_borrowed_spawn_buffers_give_the_same_result_and_refuse_reset():
    """incerto_builder_add_entity_spawner_borrowed: no host-side copy of the spawn columns; reset is refused."""
    W, H = 40, 24
    rng = np.random.default_rng(4)
    xs, ys = np.meshgrid(np.arange(W, dtype=np.int16), np.arange(H, dtype=np.int16), indexing="ij")
    host_pos = np.concatenate([np.stack([xs.ravel(), ys.ravel()], 1), np.array([[3, 3], [20, 11]], np.int16)])
    host_state = np.concatenate([np.where(rng.random(W * H) < 0.7, F.CELL_HEALTHY, F.CELL_EMPTY).astype(np.uint8),
                                 np.array([3, 3], np.uint8)])
    out = []
    for borrow in (False, True):
        b = ib.SimulationBuilder(seed=SEED)
        pos = b.component("GridPosition2D", F.I16, lanes=2)
        cell = b.component("ForestCell.state", F.U8)
        b.add_spatial_grid_2d(pos, cell, bounds=((0, 0), (W - 1, H - 1)))
        b.add_entity_spawner(lambda s: s.spawn_batch(len(host_state), {pos: host_pos, cell: host_state}))
        b.add_systems(ib.FireSpread(pos, cell, 0.6, 3), ib.BurnProgression(pos, cell), ib.Regrowth(pos, cell, 0.01))
        sim = b.build(borrow_spawn_buffers=borrow)
        sim.run(40)
        out.append(sim.read_component(cell))
        if borrow:
            with pytest.raises(ib.EngineError) as e:
                sim.reset(1)
            assert e.value.code == F.ERR_UNSUPPORTED
        else:
            sim.reset(0)
    assert np.array_equal(out[0], out[1]) and (out[0] == F.CELL_BURNED).any()


def test_run_cold_is_run_with_a_flushed_l2_before_every_step():
    """incerto_sim_run_cold (the bench's timed loop): same states, same series, same step counter as run(n); the
    device time it reports is the sum of the per-step event times."""
    W, H, steps = 96, 80, 25
    a, _, cell_a = build_device_forest(W, H, density=0.85, p_regrow=0.004, fires=5)
    b, _, cell_b = build_device_forest(W, H, density=0.85, p_regrow=0.004, fires=5)
    a.run(steps)
    b.run_cold(10)
    b.run_cold(steps - 10)
    ms, launches = b.last_run_stats()
    assert launches == steps - 10 + 1 and ms > 0   # one launch per step + the one that puts the last step on record
    assert a.step_number() == b.step_number() == steps + 1
    assert np.array_equal(a.read_component(cell_a), b.read_component(cell_b))
    ta, tb = a.get_aggregate_time_series(ib.FireStatsOf(cell_a)), b.get_aggregate_time_series(ib.FireStatsOf(cell_b))
    assert ta.len() == tb.len() == steps
    assert [stats_tuple(v) for v in ta.values()] == [stats_tuple(v) for v in tb.values()]
    o = ForestOracle(SEED, 0, W, H, density=0.85, p_regrow=0.004, fires=5).spawn().run(steps)
    assert np.array_equal(b.read_component(cell_b), o.cells()[0])


@pytest.mark.parametrize("W,H", [(24, 40), (1100, 1024)])   # the big one takes the multi-threaded host check
@pytest.mark.parametrize("borrow", [False, True])
def test_host_spawn_is_verified(W, H, borrow):
    """A host-spawned forest must be the dense x-major grid of forest_fire.rs:239-257 with valid state codes, and an
    entity outside the grid bounds is refused as the reference's SpatialGrid does (spatial_grid.rs:276-279)."""
    from incerto_b200.simulation import EngineError
    xs, ys = np.meshgrid(np.arange(W), np.arange(H), indexing="ij")
    xy = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.int16)
    st = np.full(W * H, HEALTHY, np.uint8)

    def build(xy_, st_):
        b = ib.SimulationBuilder(seed=SEED)
        pos = b.component("GridPosition2D", F.I16, lanes=2)
        cell = b.component("ForestCell.state", F.U8)
        b.add_spatial_grid_2d(pos, cell, bounds=((0, 0), (W - 1, H - 1)))
        b.add_entity_spawner(lambda s: s.spawn_batch(len(st_), {pos: xy_, cell: st_}))
        b.add_systems(ib.FireSpread(pos, cell, 0.6, 3), ib.BurnProgression(pos, cell), ib.Regrowth(pos, cell, 0.004))
        return b.build(borrow_spawn_buffers=borrow), cell

    sim, cell = build(xy, st)
    assert np.array_equal(sim.read_component(cell), st)
    swapped = xy.copy()
    k = (W * H * 3) // 4 + 5                       # two rows in the wrong order, far into the grid
    swapped[[k, k + 1]] = swapped[[k + 1, k]]
    with pytest.raises(EngineError) as e:
        build(swapped, st)
    assert e.value.code == F.ERR_UNSUPPORTED and "x-major" in str(e.value)
    bad = st.copy()
    bad[W * H - 7] = 0x41                           # not a ForestCell state code
    with pytest.raises(EngineError) as e:
        build(xy, bad)
    assert e.value.code == F.ERR_INVALID_ARGUMENT
    with pytest.raises(EngineError) as e:
        build(np.concatenate([xy, np.array([[W, 0]], np.int16)]), np.concatenate([st, np.array([3], np.uint8)]))
    assert e.value.code == F.ERR_OUT_OF_BOUNDS


@pytest.mark.parametrize("interval", [1, 3, 16, 40])
def test_fused_series_cadence_over_graph_and_single_launches(interval):
    """The FireStats record of a step is written by the next step's grid (or by the publish kernel after the last step
    of a run / of a captured graph of 16 steps): every mix of run lengths must give the series of one long run, sampled
    when step % interval == 0 (time_series.rs:68-87), and sample_aggregate must see the state after the last step."""
    W, H = 80, 70
    runs = [1, 2, 33, 5, 64, 16, 1, 47, 3]          # singles, graphs of 16 + remainders, in every order
    total = sum(runs)
    sim, _, cell = build_device_forest(W, H, density=0.85, p_regrow=0.004, fires=5, series=interval)
    o = ForestOracle(SEED, 0, W, H, density=0.85, p_regrow=0.004, fires=5).spawn()
    want = []
    done = 0
    for n in runs:
        sim.run(n)
        for _ in range(n):
            o.run(1)
            done += 1
            if done % interval == 0:
                want.append(tuple(int(v) for v in o.stats()))
        assert stats_tuple(sim.sample_aggregate(ib.FireStatsOf(cell))) == tuple(int(v) for v in o.stats())
        assert sim.step_number() == done + 1
    ts = sim.get_aggregate_time_series(ib.FireStatsOf(cell))
    assert ts.len() == total // interval
    assert [stats_tuple(v) for v in ts.values()] == want
    assert list(ts.time()) == [interval * (k + 1) for k in range(total // interval)]
    assert np.array_equal(sim.read_component(cell), o.cells()[0])
    # and again after a reset under another replica key (the publish state restarts with the spawn)
    sim.reset(5)
    o2 = ForestOracle(SEED, 5, W, H, density=0.85, p_regrow=0.004, fires=5).spawn().run(35)
    sim.run(35)
    assert stats_tuple(sim.sample_aggregate(ib.FireStatsOf(cell))) == tuple(int(v) for v in o2.stats())
    ts = sim.get_aggregate_time_series(ib.FireStatsOf(cell))
    assert ts.len() == 35 // interval
