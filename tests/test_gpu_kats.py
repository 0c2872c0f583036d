"""The reference's own tests (tests/*.rs, all RNG-free) run through the C ABI on the GPU.

Each case names the Rust test it ports; the Python facade keeps the reference's method names."""
import numpy as np
import pytest

import incerto_b200 as ib
from incerto_b200 import _ffi as F

pytestmark = pytest.mark.gpu


def test_counter():                                   # tests/test_counter.rs:34-62
    b = ib.SimulationBuilder()
    MyCounter = b.component("MyCounter", F.U64)
    b.add_systems(ib.AddConst(MyCounter, 1)).add_entity_spawner(lambda s: s.spawn({MyCounter: 0}))
    sim = b.build()
    sim.run(100)
    assert sim.sample_aggregate(ib.Sum(MyCounter)) == 100
    sim.run(100)
    assert sim.sample_aggregate(ib.Sum(MyCounter)) == 200
    assert sim.sample_single(MyCounter) == 200        # simulation.rs:79-93
    assert sim.step_number() == 201


def test_many_counters():                             # :64-93
    b = ib.SimulationBuilder()
    MyCounter = b.component("MyCounter", F.U64)
    b.add_systems(ib.AddConst(MyCounter, 1)).add_entity_spawner(lambda s: [s.spawn({MyCounter: 0}) for _ in range(56)])
    sim = b.build()
    sim.run(100)
    assert sim.sample_aggregate(ib.Sum(MyCounter)) == 100 * 56
    with pytest.raises(ib.SamplingError) as e:
        sim.sample_single(MyCounter)
    assert e.value.kind == "SingleMultipleEntities"


def test_counters_two_groups():                       # :95-152
    b = ib.SimulationBuilder()
    MyCounter = b.component("MyCounter", F.U64)
    GroupA, GroupB = b.component("GroupA"), b.component("GroupB")

    def spawn(s):
        for _ in range(50):
            s.spawn({GroupA: None, MyCounter: 0})
            s.spawn({GroupB: None, MyCounter: 0})
    b.add_systems(ib.AddConst(MyCounter, 1)).add_systems(ib.AddConst(MyCounter, 1, with_=[GroupA])).add_entity_spawner(spawn)
    sim = b.build()
    sim.run(100)
    assert sim.sample_aggregate(ib.Sum(MyCounter)) == 3 * 100 * 50
    assert sim.sample_aggregate_filtered(ib.Sum(MyCounter), ib.With(GroupA)) == 2 * 100 * 50
    assert sim.sample_aggregate_filtered(ib.Sum(MyCounter), ib.With(GroupB)) == 100 * 50
    assert sim.sample_aggregate(ib.Minimum(MyCounter)) == 100
    assert sim.sample_aggregate(ib.Maximum(MyCounter)) == 200
    assert sim.sample_aggregate(ib.Mean(MyCounter)) == 3 * 100 // 2
    assert sim.count(ib.With(GroupA)) == 50 and sim.count(ib.With(MyCounter), ib.Without(GroupA)) == 50
    assert sim.count(ib.With(GroupA), ib.With(GroupB)) == 0


def test_counter_time_series():                       # :154-193
    b = ib.SimulationBuilder()
    MyCounter = b.component("MyCounter", F.U64)
    b.add_systems(ib.AddConst(MyCounter, 1)).add_entity_spawner(lambda s: s.spawn({MyCounter: 0}))
    b.record_aggregate_time_series(ib.Sum(MyCounter), 10)
    sim = b.build()
    sim.run(100)
    ts = sim.get_aggregate_time_series(ib.Sum(MyCounter))
    assert ts.values_copied() == [10, 20, 30, 40, 50, 60, 70, 80, 90, 100]
    assert ts.time() == [10, 20, 30, 40, 50, 60, 70, 80, 90, 100] and ts.sample_interval() == 10
    assert ts.len() == 10 and ts.duration() == 100
    sim.run(10)
    assert sim.get_aggregate_time_series(ib.Sum(MyCounter)).values_copied() == [10 * k for k in range(1, 12)]
    with pytest.raises(ib.SamplingError) as e:
        sim.get_aggregate_time_series(ib.Minimum(MyCounter))
    assert e.value.kind == "TimeSeriesNotRecorded"    # simulation.rs:253


def test_counter_with_id():                           # :195-226
    b = ib.SimulationBuilder()
    MyCounter = b.component("MyCounter", F.U64)
    CounterId = b.component("CounterId", F.U64)

    def spawn(s):
        s.spawn({MyCounter: 0, CounterId: 1})
        s.spawn({MyCounter: 0, CounterId: 5})
    b.add_systems(ib.AddColumn(MyCounter, CounterId)).add_entity_spawner(spawn)
    sim = b.build()
    sim.run(100)
    assert sim.sample(MyCounter, CounterId, 1) == 100
    assert sim.sample(MyCounter, CounterId, 5) == 500
    with pytest.raises(ib.SamplingError) as e:
        sim.sample(MyCounter, CounterId, 7)
    assert e.value.kind == "EntityIdentifierNotFound"


def test_sample_identifier_not_unique_and_missing_component():
    b = ib.SimulationBuilder()
    V = b.component("V", F.U32)
    Id = b.component("Id", F.U32)
    Never = b.component("Never", F.U32)
    NeverMarker = b.component("NeverMarker")
    b.add_entity_spawner(lambda s: [s.spawn({V: 1, Id: 3}), s.spawn({V: 2, Id: 3})])
    sim = b.build()
    with pytest.raises(ib.SamplingError) as e:
        sim.sample(V, Id, 3)
    assert e.value.kind == "EntityIdentifierNotUnique"     # simulation.rs:59-62
    with pytest.raises(ib.SamplingError) as e:
        sim.sample_aggregate(ib.Sum(Never))
    assert e.value.kind == "ComponentDoesNotExist"         # simulation.rs:110-112
    with pytest.raises(ib.SamplingError) as e:
        sim.count(ib.With(NeverMarker))
    assert e.value.kind == "ComponentDoesNotExist"         # simulation.rs:166-168


def test_aggregates_int():                            # tests/test_aggregates.rs:27-57
    b = ib.SimulationBuilder()
    Item = b.component("Item", F.U64)
    b.add_entity_spawner(lambda s: [s.spawn({Item: 2 * i + 1}) for i in range(10)])
    sim = b.build()
    assert sim.sample_aggregate(ib.Minimum(Item)) == 1
    assert sim.sample_aggregate(ib.Maximum(Item)) == 19
    assert sim.sample_aggregate(ib.Mean(Item)) == 10
    assert sim.sample_aggregate(ib.Median(Item)) == 11
    assert sim.sample_aggregate(ib.Percentile(Item, 10)) == 3
    assert sim.sample_aggregate(ib.Percentile(Item, 70)) == 15
    with pytest.raises(ib.EngineError) as e:          # aggregate.rs:131-132: index == len
        sim.sample_aggregate(ib.Percentile(Item, 100))
    assert e.value.code == F.ERR_PERCENTILE_RANGE


def test_aggregates_float():                          # :59-90
    b = ib.SimulationBuilder()
    ItemFloat = b.component("ItemFloat", F.F32)
    b.add_entity_spawner(lambda s: [s.spawn({ItemFloat: 2.0 * i + 1.0}) for i in range(10)])
    sim = b.build()
    assert sim.sample_aggregate(ib.Minimum(ItemFloat)) == 1.0
    assert sim.sample_aggregate(ib.Maximum(ItemFloat)) == 19.0
    assert sim.sample_aggregate(ib.Mean(ItemFloat)) == 10.0
    assert sim.sample_aggregate(ib.Median(ItemFloat)) == 11.0
    assert sim.sample_aggregate(ib.Percentile(ItemFloat, 30)) == 7.0


@pytest.mark.parametrize("dt,npdt", [(F.U8, np.uint8), (F.I16, np.int16), (F.I32, np.int32), (F.U32, np.uint32),
                                     (F.I64, np.int64), (F.F32, np.float32), (F.F64, np.float64)])
def test_aggregates_match_oracle_on_random_columns(dt, npdt, oracle_lib):
    """Aggregators at a size the oracle finishes instantly, every dtype, with a filter."""
    import ctypes as C
    rng = np.random.default_rng(11)
    n = 100_003
    if np.issubdtype(npdt, np.floating):
        vals = (rng.standard_normal(n) * 1000).astype(npdt)
    else:
        info = np.iinfo(npdt)
        vals = rng.integers(max(info.min, -20000), min(info.max, 20000), n, endpoint=True).astype(npdt)
    keep = rng.random(n) < 0.4
    b = ib.SimulationBuilder()
    V = b.component("V", dt)
    M = b.component("M")
    b.add_entity_spawner(lambda s: s.spawn_batch(n, {V: vals, M: keep.astype(np.uint8)}))
    sim = b.build()
    name = {np.uint8: "u8", np.int16: "i16", np.int32: "i32", np.uint32: "u32", np.int64: "i64", np.float32: "f32",
            np.float64: "f64"}[npdt]
    ct = np.ctypeslib.as_ctypes_type(np.dtype(npdt))

    def oracle(arr, kind, P=0):
        out = ct()
        arr = np.ascontiguousarray(arr)
        rc = getattr(oracle_lib, f"oracle_aggregate_{name}")(arr.ctypes.data_as(C.POINTER(ct)), arr.size, kind, P,
                                                              C.byref(out))
        assert rc == 0
        return out.value
    for sub, filt in ((vals, ()), (vals[keep], (ib.With(M),)), (vals[~keep], (ib.Without(M),))):
        assert sim.sample_aggregate(ib.Minimum(V).filtered(*filt)) == oracle(sub, 3)
        assert sim.sample_aggregate(ib.Maximum(V).filtered(*filt)) == oracle(sub, 4)
        assert sim.sample_aggregate(ib.Median(V).filtered(*filt)) == oracle(sub, 6)
        for P in (0, 10, 29, 57, 99):
            assert sim.sample_aggregate(ib.Percentile(V, P).filtered(*filt)) == oracle(sub, 7, P)
        assert sim.sample_aggregate(ib.Count(V).filtered(*filt)) == sub.size
        got = sim.sample_aggregate(ib.Mean(V).filtered(*filt))
        if np.issubdtype(npdt, np.floating):
            # Mean<f32|f64> sums in arbitrary order (traits.rs:18-19): tolerance 1e-6 relative
            assert got == pytest.approx(float(np.mean(sub.astype(np.float64))), rel=1e-6, abs=1e-6)
        elif npdt in (np.int64, np.int32, np.uint32):
            s_ = int(sub.astype(object).sum())
            want = int(abs(s_) // sub.size) * (1 if s_ >= 0 else -1)   # truncating division in T
            assert got == want
    assert sim.count(ib.With(M)) == int(keep.sum()) and sim.count(ib.Without(M)) == int((~keep).sum())


def test_nan_is_not_comparable():                     # aggregate.rs:221-232
    b = ib.SimulationBuilder()
    V = b.component("V", F.F64)
    b.add_entity_spawner(lambda s: s.spawn_batch(3, {V: np.array([1.0, np.nan, 2.0])}))
    sim = b.build()
    with pytest.raises(ib.EngineError) as e:
        sim.sample_aggregate(ib.Minimum(V))
    assert e.value.code == F.ERR_NOT_COMPARABLE


def test_aggregate_no_entities_and_series_gap():      # simulation.rs:116-119, time_series.rs:79
    b = ib.SimulationBuilder()
    V = b.component("V", F.U64)
    A = b.component("A")
    b.add_systems(ib.AddConst(V, 1)).add_entity_spawner(lambda s: s.spawn({V: 0}))
    b.record_aggregate_time_series(ib.Sum(V).filtered(ib.With(A)), 1)
    sim = b.build()
    sim.run(5)
    with pytest.raises(ib.SamplingError) as e:
        sim.sample_aggregate_filtered(ib.Sum(V), ib.With(A))
    assert e.value.kind == "AggregateNoEntities"
    assert sim.get_aggregate_time_series_filtered(ib.Sum(V), ib.With(A)).len() == 0   # samples with no entities are skipped


def test_3d_spatial_grid_integration():               # tests/test_spatial_grid.rs:176-259
    b = ib.SimulationBuilder()
    Pos = b.component("GridPosition3D", F.I16, lanes=3)
    Ent = b.component("TestEntity3D", F.I32)
    b.add_spatial_grid_3d(Pos, Ent, bounds=((0, 0, 0), (4, 4, 4)))
    g = b.last_grid

    def spawn(s):
        s.spawn({Pos: (0, 0, 0), Ent: 1})
        s.spawn({Pos: (2, 2, 2), Ent: 2})
        s.spawn({Pos: (4, 4, 4), Ent: 3})
        s.spawn({Pos: (1, 2, 3), Ent: 4})
    b.add_entity_spawner(spawn)
    sim = b.build()
    assert sim.grid_num_entities(g) == 0              # populated by PreUpdate of the first step
    sim.run(1)
    near = []
    for dx in range(-2, 3):
        for dy in range(-2, 3):
            for dz in range(-2, 3):
                if abs(dx) + abs(dy) + abs(dz) <= 2:
                    near += sim.grid_entities_at(g, (2 + dx, 2 + dy, 2 + dz))
    assert sorted(near) == [1, 3]
    assert sim.sample_aggregate(ib.Count(Ent)) == 4
    assert sim.grid_num_entities(g) == 4
    assert sim.grid_neighbors_of(g, (1, 1, 1)) == [0, 1]          # 26-neighbourhood of (1,1,1): (0,0,0) and (2,2,2)
    assert sim.grid_neighbors_of(g, (2, 2, 3), orthogonal=True) == [1, 3]
    assert sim.grid_is_empty(g, (3, 3, 3)) and not sim.grid_is_empty(g, (4, 4, 4))
    assert sim.grid_position_of(g, 3, dim=3) == (1, 2, 3)
    assert sim.grid_in_bounds(g, (4, 4, 4)) and not sim.grid_in_bounds(g, (5, 0, 0))   # inclusive bounds (:143-174)


def test_spatial_grid_survives_two_steps():           # :261-299
    b = ib.SimulationBuilder()
    Pos = b.component("GridPosition2D", F.I16, lanes=2)
    Ent = b.component("TestResetEntity", F.I32)
    b.add_spatial_grid_2d(Pos, Ent, bounds=((0, 0), (4, 4)))
    g = b.last_grid
    b.add_entity_spawner(lambda s: [s.spawn({Pos: p, Ent: v}) for p, v in (((0, 0), 1), ((2, 2), 2), ((4, 4), 3))])
    sim = b.build()
    sim.run(2)
    assert sim.sample_aggregate(ib.Count(Ent)) == 3
    assert sim.grid_entities_at(g, (2, 2)) == [1]
    # Moore / von-Neumann neighbour queries honour the bounds (spatial_grid.rs:331-372)
    assert sim.grid_neighbors_of(g, (1, 1)) == [0, 1]
    assert sim.grid_neighbors_of(g, (1, 1), orthogonal=True) == []
    assert sim.grid_neighbors_of(g, (4, 3), orthogonal=True) == [2]


def test_grid_out_of_bounds_is_an_error():            # spatial_grid.rs:276-279 panics
    b = ib.SimulationBuilder()
    Pos = b.component("GridPosition2D", F.I16, lanes=2)
    Ent = b.component("E", F.I32)
    b.add_spatial_grid_2d(Pos, Ent, bounds=((0, 0), (4, 4)))
    b.add_entity_spawner(lambda s: [s.spawn({Pos: (5, 0), Ent: 1})])
    sim = b.build()
    with pytest.raises(ib.EngineError) as e:
        sim.run(1)
    assert e.value.code == F.ERR_OUT_OF_BOUNDS


def test_grid_matches_oracle_on_random_population(oracle_lib):
    import ctypes as C
    rng = np.random.default_rng(5)
    n, G = 20_000, 37
    pos = rng.integers(0, G, (n, 2)).astype(np.int16)
    member = rng.random(n) < 0.7
    b = ib.SimulationBuilder()
    Pos = b.component("GridPosition2D", F.I16, lanes=2)
    Val = b.component("Val", F.U32)
    Person = b.component("Person")
    b.add_spatial_grid_2d(Pos, Person, bounds=((0, 0), (G - 1, G - 1)))
    g = b.last_grid
    b.add_entity_spawner(lambda s: s.spawn_batch(n, {Pos: pos, Val: np.arange(n), Person: member.astype(np.uint8)}))
    sim = b.build()
    sim.run(1)
    og = oracle_lib.oracle_grid_new(2, 1, (C.c_int32 * 4)(0, 0, G - 1, G - 1))
    for e in np.nonzero(member)[0]:
        oracle_lib.oracle_grid_insert(og, int(e), (C.c_int32 * 2)(int(pos[e, 0]), int(pos[e, 1])))
    assert sim.grid_num_entities(g) == oracle_lib.oracle_grid_num_entities(og) == int(member.sum())
    out = (C.c_uint32 * n)()
    for (x, y) in [(0, 0), (G - 1, G - 1), (5, 7), (18, 18), (0, 20), (36, 1)]:
        q = (C.c_int32 * 2)(x, y)
        for mode, fn in ((0, lambda: sim.grid_entities_at(g, (x, y))), (1, lambda: sim.grid_neighbors_of(g, (x, y))),
                         (2, lambda: sim.grid_neighbors_of(g, (x, y), orthogonal=True))):
            k = oracle_lib.oracle_grid_query(og, mode, q, out, n)
            assert fn() == sorted(out[i] for i in range(k))
    oracle_lib.oracle_grid_free(og)


def test_per_entity_time_series():                    # record_time_series / get_time_series (simulation.rs:186-216)
    b = ib.SimulationBuilder()
    V = b.component("V", F.U64)
    Id = b.component("Id", F.U64)
    b.add_systems(ib.AddColumn(V, Id)).add_entity_spawner(lambda s: [s.spawn({V: 0, Id: i}) for i in (1, 2, 3)])
    b.record_time_series(V, Id, 2)
    with pytest.raises(ib.BuilderError):
        b.record_time_series(V, Id, 3)
    sim = b.build()
    sim.run(7)
    ts = sim.get_time_series(V, Id, 3)
    assert ts.time() == [2, 4, 6] and ts.values() == [6, 12, 18] and ts.duration() == 6
    with pytest.raises(ib.SamplingError) as e:
        sim.get_time_series(V, Id, 9)
    assert e.value.kind == "EntityIdentifierNotFound"


def test_global_aggregate_needs_a_communicator():
    """incerto_sim_sample_aggregate_global is a collective over the attached communicator (checked against numpy on
    several GPUs by tools/check_global_aggregates.py); without one it says so."""
    b = ib.SimulationBuilder()
    v = b.component("v", F.I32)
    b.add_entity_spawner(lambda s: s.spawn_batch(4, {v: np.array([3, 1, 4, 1], np.int32)}))
    sim = b.build()
    assert sim.sample_aggregate(ib.Median(v)) == 3
    with pytest.raises(ib.EngineError) as e:
        sim.sample_aggregate_global(ib.Median(v))
    assert e.value.code == F.ERR_COMM


def test_unbounded_spatial_grid_matches_oracle(oracle_lib):
    """add_spatial_grid(None) (simulation_builder.rs:114-121): no position is out of bounds (spatial_grid.rs:265-270,
    negative coordinates included), neighbour queries are not filtered by bounds; the device index is built on demand
    over the bounding box of the entities."""
    import ctypes as C
    rng = np.random.default_rng(6)
    n = 5000
    pos = rng.integers(-40, 41, (n, 2)).astype(np.int16)
    pos[0] = (-300, 250)                               # an outlier widens the bounding box
    b = ib.SimulationBuilder()
    Pos = b.component("GridPosition2D", F.I16, lanes=2)
    Val = b.component("Val", F.U32)
    b.add_spatial_grid_2d(Pos, Val, bounds=None)
    g = b.last_grid
    b.add_entity_spawner(lambda s: s.spawn_batch(n, {Pos: pos, Val: np.arange(n)}))
    sim = b.build()
    sim.run(1)
    og = oracle_lib.oracle_grid_new(2, 0, None)
    for e in range(n):
        oracle_lib.oracle_grid_insert(og, e, (C.c_int32 * 2)(int(pos[e, 0]), int(pos[e, 1])))
    assert sim.grid_num_entities(g) == oracle_lib.oracle_grid_num_entities(og) == n
    out = (C.c_uint32 * n)()
    for (x, y) in [(0, 0), (-40, -40), (40, 40), (-300, 250), (-301, 250), (41, 0), (5000, 5000), (-7, 13)]:
        q = (C.c_int32 * 2)(x, y)
        assert sim.grid_in_bounds(g, (x, y))
        for mode, fn in ((0, lambda: sim.grid_entities_at(g, (x, y))), (1, lambda: sim.grid_neighbors_of(g, (x, y))),
                         (2, lambda: sim.grid_neighbors_of(g, (x, y), orthogonal=True))):
            k = oracle_lib.oracle_grid_query(og, mode, q, out, n)
            assert fn() == sorted(out[i] for i in range(k)), (x, y, mode)
    oracle_lib.oracle_grid_free(og)


def test_run_refuses_to_wrap_the_32_bit_step_counter():
    """StepNumber and the Philox counter word are 32 bit on the device: a run past 2^32 - 1 is an error, not a
    silent repeat of the draws."""
    b = ib.SimulationBuilder()
    c = b.component("Counter", F.I32)
    b.add_entity_spawner(lambda s: s.spawn({c: 0}))
    b.add_systems(ib.AddConst(c, 1))
    sim = b.build()
    sim.run(3)
    with pytest.raises(ib.EngineError) as e:
        sim.run(1 << 32)
    assert e.value.code == F.ERR_INVALID_ARGUMENT
    sim.run(2)
    assert sim.sample_single(c) == 5


@pytest.mark.parametrize("dt,npdt", [(F.F64, np.float64), (F.I32, np.int32), (F.U16, np.uint16)])
def test_median_and_percentile_across_archetype_tables(dt, npdt):
    """sample_aggregate queries every entity that has the component, whatever else it carries (simulation.rs:107-121):
    Median / Percentile over a component living in three archetype tables = one selection over all of them
    (aggregate.rs:107-155: sorted[len / 2], sorted[floor(P / 100 * len)])."""
    rng = np.random.default_rng(12)
    sizes = (4001, 1, 2500)
    if npdt is np.float64:
        vals = [rng.normal(size=n).astype(npdt) for n in sizes]
    else:
        vals = [rng.integers(0, 30000, n).astype(npdt) for n in sizes]
    b = ib.SimulationBuilder()
    Val = b.component("Val", dt)
    A, B = b.component("ExtraA", F.U32), b.component("ExtraB", F.F64)
    Tag = b.component("Tag")
    tags = [(rng.random(n) < 0.5).astype(np.uint8) for n in sizes]
    b.add_entity_spawner(lambda s: s.spawn_batch(sizes[0], {Val: vals[0], Tag: tags[0]}))
    b.add_entity_spawner(lambda s: s.spawn_batch(sizes[1], {Val: vals[1], A: np.zeros(sizes[1], np.uint32), Tag: tags[1]}))
    b.add_entity_spawner(lambda s: s.spawn_batch(sizes[2], {Val: vals[2], A: np.zeros(sizes[2], np.uint32),
                                                            B: np.zeros(sizes[2]), Tag: tags[2]}))
    sim = b.build()
    allv = np.sort(np.concatenate(vals))
    assert sim.sample_aggregate(ib.Count(Val)) == len(allv)
    assert sim.sample_aggregate(ib.Median(Val)) == allv[len(allv) // 2]
    for p in (0, 1, 29, 50, 57, 99):
        assert sim.sample_aggregate(ib.Percentile(Val, p)) == allv[int(np.floor(p / 100.0 * len(allv)))]
    with pytest.raises(ib.EngineError) as e:
        sim.sample_aggregate(ib.Percentile(Val, 100))
    assert e.value.code == F.ERR_PERCENTILE_RANGE
    sel = np.sort(np.concatenate([v[t.astype(bool)] for v, t in zip(vals, tags)]))
    assert sim.sample_aggregate_filtered(ib.Median(Val), ib.With(Tag)) == sel[len(sel) // 2]
    assert sim.sample_aggregate(ib.Minimum(Val)) == allv[0] and sim.sample_aggregate(ib.Maximum(Val)) == allv[-1]
