"""Pins the oracle against every known-answer test the reference holds (its tests/*.rs are all RNG-free).

Each test names the reference test it ports.  Two oracle layers are checked: the pure-Python library
restatement (oracle/incerto_ref.py, runs arbitrary systems like the reference) and the C++ pieces the
workload oracles use (aggregators, SpatialGrid)."""
import ctypes as C

import numpy as np
import pytest

from oracle import incerto_ref as R


# components of tests/test_counter.rs:5-15
class MyCounter:
    def __init__(self, v=0): self.v = v


class GroupA: pass
class GroupB: pass


class CounterId:
    def __init__(self, v): self.v = v
    def __eq__(self, o): return isinstance(o, CounterId) and self.v == o.v
    def __hash__(self): return hash(self.v)


def counter_sum(cs): return sum(c.v for c in cs)     # test_counter.rs:17-24


def test_counter():                                   # test_counter.rs:34-62
    def system(ctx):
        q = ctx.world.query(MyCounter)
        assert len(q) == 1
        q[0][1].v += 1
    sim = R.SimulationBuilder().add_systems(system).add_entity_spawner(lambda s: s.spawn(MyCounter(0))).build()
    sim.run(100)
    assert sim.sample_aggregate(MyCounter, counter_sum) == 100
    sim.run(100)
    assert sim.sample_aggregate(MyCounter, counter_sum) == 200


def test_many_counters():                             # :64-93
    def system(ctx):
        for _, c in ctx.world.query(MyCounter):
            c.v += 1
    sim = R.SimulationBuilder().add_systems(system).add_entity_spawner(
        lambda s: [s.spawn(MyCounter(0)) for _ in range(56)]).build()
    sim.run(100)
    assert sim.sample_aggregate(MyCounter, counter_sum) == 5600


def test_counters_two_groups():                       # :95-152
    def inc_all(ctx):
        for _, c in ctx.world.query(MyCounter):
            c.v += 1

    def inc_a(ctx):
        for _, c in ctx.world.query(MyCounter, with_=(GroupA,)):
            c.v += 1

    def spawn(s):
        for _ in range(50):
            s.spawn(GroupA(), MyCounter(0))
            s.spawn(GroupB(), MyCounter(0))
    sim = R.SimulationBuilder().add_systems(inc_all).add_systems(inc_a).add_entity_spawner(spawn).build()
    sim.run(100)
    assert sim.sample_aggregate(MyCounter, counter_sum) == 3 * 100 * 50
    assert sim.sample_aggregate(MyCounter, counter_sum, with_=(GroupA,)) == 2 * 100 * 50
    assert sim.sample_aggregate(MyCounter, counter_sum, with_=(GroupB,)) == 100 * 50
    vals = lambda cs: [c.v for c in cs]
    assert sim.sample_aggregate(MyCounter, lambda cs: R.minimum(vals(cs))) == 100
    assert sim.sample_aggregate(MyCounter, lambda cs: R.maximum(vals(cs))) == 200
    assert sim.sample_aggregate(MyCounter, lambda cs: R.mean(vals(cs), integer=True)) == 150


def test_counter_time_series():                       # :154-193
    def system(ctx):
        ctx.world.query(MyCounter)[0][1].v += 1
    b = R.SimulationBuilder().add_systems(system).add_entity_spawner(lambda s: s.spawn(MyCounter(0)))
    b.record_aggregate_time_series(MyCounter, "usize", counter_sum, 10)
    sim = b.build()
    sim.run(100)
    ts = sim.get_aggregate_time_series(MyCounter, "usize")
    assert ts.values_copied() == [10, 20, 30, 40, 50, 60, 70, 80, 90, 100]
    assert ts.time() == [10, 20, 30, 40, 50, 60, 70, 80, 90, 100]
    sim.run(10)
    assert sim.get_aggregate_time_series(MyCounter, "usize").values_copied() == [10 * k for k in range(1, 12)]


def test_counter_with_id():                           # :195-226
    def system(ctx):
        for _, i, c in ctx.world.query(CounterId, MyCounter):
            c.v += i.v

    def spawn(s):
        s.spawn(MyCounter(0), CounterId(1))
        s.spawn(MyCounter(0), CounterId(5))
    sim = R.SimulationBuilder().add_systems(system).add_entity_spawner(spawn).build()
    sim.run(100)
    assert sim.sample(MyCounter, CounterId, CounterId(1), lambda c: c.v) == 100
    assert sim.sample(MyCounter, CounterId, CounterId(5), lambda c: c.v) == 500
    with pytest.raises(R.SamplingError) as e:
        sim.sample(MyCounter, CounterId, CounterId(7), lambda c: c.v)
    assert e.value.kind == "EntityIdentifierNotFound"


def test_aggregates_int_and_float():                  # test_aggregates.rs:27-90
    ints = [2 * i + 1 for i in range(10)]
    assert R.minimum(ints) == 1 and R.maximum(ints) == 19
    assert R.mean(ints, integer=True) == 10
    assert R.median(ints) == 11
    assert R.percentile(ints, 10) == 3 and R.percentile(ints, 70) == 15
    fl = [2.0 * i + 1.0 for i in range(10)]
    assert R.minimum(fl) == 1.0 and R.maximum(fl) == 19.0 and R.mean(fl, integer=False) == 10.0
    assert R.median(fl) == 11.0 and R.percentile(fl, 30) == 7.0
    with pytest.raises(IndexError):                   # aggregate.rs:131-132 with P = 100
        R.percentile(ints, 100)
    with pytest.raises(ValueError):                   # NaN panics, aggregate.rs:227-231
        R.minimum([1.0, float("nan")])


AGG = {"min": 3, "max": 4, "mean": 5, "median": 6, "pct": 7}


def _cpp_agg(lib, arr, kind, P=0):
    arr = np.ascontiguousarray(arr)
    name = {np.dtype(np.uint64): "u64", np.dtype(np.float32): "f32", np.dtype(np.int32): "i32",
            np.dtype(np.float64): "f64", np.dtype(np.uint8): "u8"}[arr.dtype]
    ct = np.ctypeslib.as_ctypes_type(arr.dtype)
    out = ct()
    rc = getattr(lib, f"oracle_aggregate_{name}")(arr.ctypes.data_as(C.POINTER(ct)), arr.size, kind, P, C.byref(out))
    return rc, out.value


def test_cpp_aggregators_known_answers(oracle_lib):   # test_aggregates.rs:27-90 on the C++ layer
    ints = np.array([2 * i + 1 for i in range(10)], np.uint64)
    assert _cpp_agg(oracle_lib, ints, AGG["min"]) == (0, 1)
    assert _cpp_agg(oracle_lib, ints, AGG["max"]) == (0, 19)
    assert _cpp_agg(oracle_lib, ints, AGG["mean"]) == (0, 10)
    assert _cpp_agg(oracle_lib, ints, AGG["median"]) == (0, 11)
    assert _cpp_agg(oracle_lib, ints, AGG["pct"], 10) == (0, 3)
    assert _cpp_agg(oracle_lib, ints, AGG["pct"], 70) == (0, 15)
    fl = np.array([2.0 * i + 1.0 for i in range(10)], np.float32)
    assert _cpp_agg(oracle_lib, fl, AGG["median"]) == (0, 11.0)
    assert _cpp_agg(oracle_lib, fl, AGG["pct"], 30) == (0, 7.0)
    assert _cpp_agg(oracle_lib, fl, AGG["mean"]) == (0, 10.0)
    assert _cpp_agg(oracle_lib, ints, AGG["pct"], 100)[0] == 36
    assert _cpp_agg(oracle_lib, np.array([1.0, np.nan], np.float32), AGG["min"])[0] == 35
    # Percentile index is floor((P/100) * len) in f64, not P*len/100
    v = np.arange(100, dtype=np.uint64)
    for P, want in ((29, 28), (57, 56), (58, 57), (50, 50), (0, 0), (99, 99)):
        assert _cpp_agg(oracle_lib, v, AGG["pct"], P) == (0, want)


def test_record_time_series_conflicts():              # test_builder.rs:27-72
    class MyValue: pass
    class MyMarker: pass
    b = R.SimulationBuilder()
    b.record_aggregate_time_series(MyValue, "usize", counter_sum, 8)
    with pytest.raises(R.BuilderError) as e:
        b.record_aggregate_time_series(MyValue, "usize", counter_sum, 12)
    assert e.value.kind == "TimeSeriesRecordingConflict"
    b.record_aggregate_time_series(MyValue, "bool", lambda cs: not cs, 8)
    b2 = R.SimulationBuilder()
    b2.record_aggregate_time_series(MyValue, "usize", counter_sum, 8, with_=(MyMarker,))
    b2.record_aggregate_time_series(MyValue, "usize", counter_sum, 8)


def test_grid_position_neighbors():                   # test_spatial_grid.rs:8-61
    n = [p.coords for p in R.GridPosition(1, 1).neighbors()]
    assert len(n) == 8 and set(n) == {(0, 0), (1, 0), (2, 0), (0, 1), (2, 1), (0, 2), (1, 2), (2, 2)}
    assert n == [(0, 0), (1, 0), (2, 0), (0, 1), (2, 1), (0, 2), (1, 2), (2, 2)]   # NW N NE W E SW S SE
    o = [p.coords for p in R.GridPosition(1, 1).neighbors_orthogonal()]
    assert o == [(1, 0), (0, 1), (2, 1), (1, 2)]     # N W E S (spatial_grid.rs:67)


def test_grid_bounds_inclusive():                     # :63-94 and :143-174
    b = R.GridBounds((0, 0), (9, 9))
    for p in ((0, 0), (9, 9), (5, 5)):
        assert b.contains(p)
    for p in ((-1, 0), (0, -1), (10, 5), (5, 10)):
        assert not b.contains(p)
    b3 = R.GridBounds((0, 0, 0), (9, 9, 9))
    assert b3.contains((0, 0, 0)) and b3.contains((9, 9, 9)) and b3.contains((5, 5, 5))
    for p in ((-1, 0, 0), (0, -1, 0), (0, 0, -1), (10, 5, 5), (5, 10, 5), (5, 5, 10)):
        assert not b3.contains(p)


def test_3d_neighbors():                              # :96-141
    n = [p.coords for p in R.GridPosition(1, 1, 1).neighbors()]
    assert len(n) == 26 and (1, 1, 1) not in n
    for p in ((0, 0, 0), (2, 2, 2), (1, 1, 0), (1, 1, 2)):
        assert p in n
    o = [p.coords for p in R.GridPosition(1, 1, 1).neighbors_orthogonal()]
    assert o == [(0, 1, 1), (2, 1, 1), (1, 0, 1), (1, 2, 1), (1, 1, 0), (1, 1, 2)]   # W E N S DOWN UP (:101)


def test_cpp_neighbor_orders_match_python(oracle_lib):
    buf = (C.c_int32 * (26 * 3))()
    for dim, orth, ref in ((2, 0, R.GridPosition(0, 0).neighbors()), (2, 1, R.GridPosition(0, 0).neighbors_orthogonal()),
                           (3, 0, R.GridPosition(0, 0, 0).neighbors()), (3, 1, R.GridPosition(0, 0, 0).neighbors_orthogonal())):
        n = oracle_lib.oracle_neighbor_offsets(dim, orth, buf)
        got = [tuple(buf[3 * i + k] for k in range(dim)) for i in range(n)]
        assert got == [p.coords for p in ref]


class GridEntity:
    def __init__(self, v): self.v = v


def test_spatial_grid_integration_2d_and_3d():        # :176-259 (3-D) and the 2-D analogue
    seen = {}

    def system(ctx):
        grid = ctx.res((R.SpatialGrid, GridEntity))
        near = []
        for dx in range(-2, 3):
            for dy in range(-2, 3):
                for dz in range(-2, 3):
                    if abs(dx) + abs(dy) + abs(dz) <= 2:
                        near.extend(grid.entities_at(R.GridPosition(2 + dx, 2 + dy, 2 + dz)))
        seen["near"] = near
        seen["n"] = grid.num_entities()

    def spawn(s):
        s.spawn(R.GridPosition(0, 0, 0), GridEntity(1))
        s.spawn(R.GridPosition(2, 2, 2), GridEntity(2))
        s.spawn(R.GridPosition(4, 4, 4), GridEntity(3))
        s.spawn(R.GridPosition(1, 2, 3), GridEntity(4))
    sim = (R.SimulationBuilder().add_spatial_grid(GridEntity, R.GridBounds((0, 0, 0), (4, 4, 4)))
           .add_entity_spawner(spawn).add_systems(system).build())
    sim.run(1)
    assert seen["n"] == 4 and sorted(seen["near"]) == [1, 3]     # (2,2,2) itself and (1,2,3) at distance 2
    assert sim.sample_aggregate(GridEntity, len) == 4


def test_spatial_grid_survives_two_steps():           # :261-299
    def spawn(s):
        for p, v in (((0, 0), 1), ((2, 2), 2), ((4, 4), 3)):
            s.spawn(R.GridPosition(*p), GridEntity(v))
    sim = R.SimulationBuilder().add_spatial_grid(GridEntity, R.GridBounds((0, 0), (4, 4))).add_entity_spawner(spawn).build()
    sim.run(2)
    assert sim.sample_aggregate(GridEntity, len) == 3


def test_grid_snapshot_and_deferred_commands():
    """SURVEY §3.3(2),(4): commands apply at the end of Update; the grid sees positions of the previous step."""
    log = []

    def mover(ctx):
        for e, p in ctx.world.query(R.GridPosition):
            p.coords = (p.coords[0] + 1, p.coords[1])

    def observer(ctx):
        grid = ctx.res((R.SpatialGrid, GridEntity))
        log.append(sorted(grid.entity_to_position[e].coords for e in grid.entity_to_position))

    def killer(ctx):
        if ctx.res(R.StepNumber).value == 2:
            ctx.commands.despawn(0)

    sim = (R.SimulationBuilder().add_spatial_grid(GridEntity, R.GridBounds((0, 0), (9, 9)))
           .add_entity_spawner(lambda s: [s.spawn(R.GridPosition(0, 0), GridEntity(1)), s.spawn(R.GridPosition(5, 5), GridEntity(2))])
           .add_systems(mover, observer, killer).build())
    sim.run(3)
    assert log[0] == [(0, 0), (5, 5)]          # step 1 sees the spawn positions
    assert log[1] == [(1, 0), (6, 5)]          # step 2 sees the positions after step 1
    assert log[2] == [(7, 5)]                  # entity 0 despawned at the end of step 2, cleaned up in PreUpdate(3)


def test_cpp_grid_matches_python_grid(oracle_lib):
    rng = np.random.default_rng(3)
    bounds = (C.c_int32 * 4)(0, 0, 7, 7)
    g = oracle_lib.oracle_grid_new(2, 1, bounds)
    py = R.SpatialGrid(R.GridBounds((0, 0), (7, 7)))
    for e in range(200):
        p = [int(x) for x in rng.integers(0, 8, 2)]
        assert oracle_lib.oracle_grid_insert(g, e, (C.c_int32 * 2)(*p)) == 0
        py.insert_or_update(e, R.GridPosition(*p))
    for e in range(0, 200, 3):       # moves
        p = [int(x) for x in rng.integers(0, 8, 2)]
        oracle_lib.oracle_grid_insert(g, e, (C.c_int32 * 2)(*p))
        py.insert_or_update(e, R.GridPosition(*p))
    for e in range(0, 200, 7):       # removals
        oracle_lib.oracle_grid_remove(g, e)
        py.remove(e)
    assert oracle_lib.oracle_grid_insert(g, 999, (C.c_int32 * 2)(8, 0)) == 34     # outside bounds panics (:276-279)
    out = (C.c_uint32 * 1024)()
    assert oracle_lib.oracle_grid_num_entities(g) == py.num_entities()
    for x in range(8):
        for y in range(8):
            q = (C.c_int32 * 2)(x, y)
            for mode, fn in ((0, py.entities_at), (1, py.neighbors_of), (2, py.orthogonal_neighbors_of)):
                n = oracle_lib.oracle_grid_query(g, mode, q, out, 1024)
                assert sorted(out[i] for i in range(n)) == sorted(fn(R.GridPosition(x, y)))
    oracle_lib.oracle_grid_free(g)
