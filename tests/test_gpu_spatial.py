"""pandemic_spatial (examples/pandemic_spatial.rs) on the GPU vs the CPU oracle: bit-exact positions, disease
states, quarantine counters, contact histories, survivors and the PandemicStats series."""
import numpy as np
import pytest

import incerto_b200 as ib
from incerto_b200 import _ffi as F
from conftest import SEED
from oracle.spatial import SpatialConfig, SpatialOracle

pytestmark = pytest.mark.gpu

ALL = 0x7f
ORDER = [(64, ib.PeopleMoveWithSocialDistancing), (1, ib.DiseaseIncubationProgression), (2, ib.SpatialDiseaseTransmission),
         (4, ib.DiseaseRecoveryAndDeath), (8, ib.UpdateContactHistory), (16, ib.ProcessContactTracing),
         (32, ib.UpdateQuarantineStatus)]          # registration order of the example (:173-189)


def build(cfg: SpatialConfig, replica=0, host=None, series=True):
    b = ib.SimulationBuilder(seed=SEED, replica=replica)
    position = b.component("GridPosition2D", F.I16, lanes=2)
    disease = b.component("Person.disease_state", F.U8)
    social = b.component("Person.social_distancing")
    quarantined = b.component("Quarantined.remaining_duration", F.U8)
    history = b.component("ContactHistory", F.U32)
    G = cfg.grid_size
    sc = ib.SpatialComponents(position, disease, social, quarantined, history, grid_size=G,
                              infection_radius=cfg.infection_radius, chance_infect_distance_1=cfg.chance_infect_1,
                              chance_infect_distance_2=cfg.chance_infect_2, chance_recover=cfg.chance_recover,
                              chance_die=cfg.chance_die, incubation_period=cfg.incubation_period,
                              crowding_threshold=cfg.crowding_threshold,
                              contact_quarantine_duration=cfg.quarantine_duration, quarantine_center=cfg.zone(),
                              quarantine_radius=cfg.zone_radius, contact_history_clear_above=cfg.history_clear_above,
                              chance_attempt_move=cfg.chance_attempt_move)
    b.add_spatial_grid_2d(position, disease, bounds=((0, 0), (G - 1, G - 1)))      # :166-172
    if host is None:
        b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_SPATIAL_POPULATION, F.SpawnSpatialParams(
            position.handle, disease.handle, social.handle, quarantined.handle, history.handle, cfg.population, G,
            cfg.social_compliance, cfg.chance_start_infected, cfg.incubation_period)))
    else:
        xy, d, sd = host
        b.add_entity_spawner(lambda s: s.spawn_batch(len(d), {position: xy, disease: d, social: np.asarray(sd, np.uint8),
                                                              quarantined: np.zeros(len(d), np.uint8), history: None}))
    b.add_systems(*[mk(sc) for bit, mk in ORDER if cfg.systems & bit])
    stats = ib.PandemicStatsOf(disease, cfg.population if host is None else len(host[1]))
    if series:
        b.record_aggregate_time_series(stats, 1)                                   # :191
    sim = b.build()
    return sim, sc, stats


def astuple(ps):
    return (ps.healthy_count, ps.exposed_count, ps.infectious_count, ps.recovered_count, ps.dead_count, ps.total_population)


def check(sim, sc, stats, o, series=True, history=True):
    st = o.state()
    alive = st["alive"]
    xy, uids = sim.read_component(sc.position, with_uids=True)
    assert np.array_equal(uids, np.nonzero(alive)[0])                  # the same people survived
    assert np.array_equal(xy.astype(np.int32), st["xy"][alive])
    assert np.array_equal(sim.read_component(sc.disease_state), st["disease"][alive])
    assert np.array_equal(sim.read_component(sc.quarantined).astype(np.uint32), st["quarantined"][alive])
    assert np.array_equal(sim.read_component(sc.social_distancing, rows_of=sc.position).astype(bool), st["social"][alive])
    if history:
        got = sim.read_contact_history(sc.contact_history, sc.position)
        want = [set((int(k) & 0xffff, int(k) >> 16) for k in row[1:1 + row[0]]) for row in st["history"][alive]]
        assert got == want
    if alive.any():
        assert astuple(sim.sample_aggregate(stats)) == o.stats()
    if series:
        ts = sim.get_aggregate_time_series(stats)
        want = o.series()
        assert len(ts) == len(want)
        assert np.array_equal(np.array([astuple(v) for v in ts.values()], dtype=np.uint64).reshape(-1, 6), want)


def test_reference_config():
    """examples/pandemic_spatial.rs:28-58 as-is: 2000 people on 40x40 for its 500 steps."""
    cfg = SpatialConfig()
    sim, sc, stats = build(cfg)
    o = SpatialOracle(SEED, 0, cfg).spawn()
    check(sim, sc, stats, o)                                            # spawn parity (:279-313)
    for chunk in (1, 1, 1, 17, 80, 400):
        sim.run(chunk)
        o.run(chunk)
        check(sim, sc, stats, o)
    s = o.stats()
    assert s[4] > 0 and s[3] > 0                                        # deaths and recoveries happened


@pytest.mark.parametrize("n,G,replica", [(1, 1, 0), (2, 1, 1), (50, 3, 2), (4097, 64, 3), (30_000, 150, 4), (60_000, 120, 5)])
def test_parity_sizes(n, G, replica):
    cfg = SpatialConfig(population=n, grid_size=G, chance_start_infected=0.05, chance_die=0.01, zone_radius=min(8, G // 4))
    sim, sc, stats = build(cfg, replica=replica)
    o = SpatialOracle(SEED, replica, cfg).spawn()
    for chunk in (1, 2, 17, 60):
        sim.run(chunk)
        o.run(chunk)
        check(sim, sc, stats, o)


@pytest.mark.parametrize("systems", [1, 2 | 1, 4 | 1, 8, 8 | 64, 16 | 8 | 64, 32 | 16 | 8, 64, 64 | 2 | 1, ALL & ~16, ALL & ~64, ALL & ~4])
def test_system_subsets(systems):
    cfg = SpatialConfig(population=3000, grid_size=30, chance_start_infected=0.1, chance_die=0.02, zone_radius=4, systems=systems)
    sim, sc, stats = build(cfg)
    o = SpatialOracle(SEED, 0, cfg).spawn()
    for chunk in (1, 40):
        sim.run(chunk)
        o.run(chunk)
        check(sim, sc, stats, o)


def test_without_tracing_people_keep_moving_and_histories_cycle():
    """No contact tracing: nobody is quarantined, histories reach 20 entries and are cleared by the 21st (:549-559)."""
    cfg = SpatialConfig(population=1500, grid_size=50, systems=ALL & ~16, chance_attempt_move=0.9, zone_radius=3)
    sim, sc, stats = build(cfg)
    o = SpatialOracle(SEED, 0, cfg).spawn()
    seen_full = False
    for _ in range(12):
        sim.run(5)
        o.run(5)
        check(sim, sc, stats, o)
        lens = o.state()["history"][:, 0]
        seen_full = seen_full or lens.max() >= 19
    assert seen_full and (o.state()["quarantined"] == 0).all()


def test_crowded_cells_and_walls_host_spawn():
    """Everybody packed into a few cells: occupancy above CROWDING_THRESHOLD blocks moves (:403-408), long source
    lists per cell, movers at the walls lose candidates (:348-355)."""
    n, G = 900, 8
    rng = np.random.default_rng(3)
    xy = rng.choice([0, 1, G - 1], size=(n, 2)).astype(np.int16)
    d = np.where(rng.random(n) < 0.2, F.DISEASE_INFECTIOUS, np.where(rng.random(n) < 0.1, 3, F.DISEASE_HEALTHY)).astype(np.uint8)
    sd = rng.random(n) < 0.5
    cfg = SpatialConfig(population=n, grid_size=G, zone_radius=1, chance_die=0.01, systems=ALL & ~16)
    sim, sc, stats = build(cfg, host=(xy, d, sd))
    o = SpatialOracle(SEED, 0, cfg).spawn_explicit(xy, d, sd)
    for _ in range(25):
        sim.run(1)
        o.run(1)
        check(sim, sc, stats, o)


def test_tag_wraparound_and_reset():
    """Source-list heads carry a 6-bit step tag: run across several wrap-arounds; then reset to another replica."""
    cfg = SpatialConfig(population=2500, grid_size=45, chance_recover=0.002, chance_die=0.0005)
    sim, sc, stats = build(cfg)
    o = SpatialOracle(SEED, 0, cfg).spawn()
    sim.run(300)
    o.run(300)
    check(sim, sc, stats, o)
    sim.reset(9)
    o9 = SpatialOracle(SEED, 9, cfg).spawn()
    check(sim, sc, stats, o9, series=False)
    sim.run(30)
    o9.run(30)
    check(sim, sc, stats, o9)


def test_out_of_bounds_person_is_an_error():
    cfg = SpatialConfig(population=1, grid_size=10)
    sim, sc, stats = build(cfg, host=(np.array([[10, 3]], np.int16), np.array([0], np.uint8), [False]))
    with pytest.raises(ib.EngineError) as e:
        sim.run(1)
    assert e.value.code == F.ERR_OUT_OF_BOUNDS


def test_host_grid_queries():
    """SpatialGrid<IVec2, Person> views after run(k): the PreUpdate snapshot of step k."""
    cfg = SpatialConfig(population=600, grid_size=12, zone_radius=2)
    sim, sc, stats = build(cfg)
    o = SpatialOracle(SEED, 0, cfg).spawn()
    sim.run(3)
    o.run(2)
    snap = o.state()
    for cell in [(0, 0), (5, 6), (11, 11)]:
        want = sorted(np.nonzero(snap["alive"] & (snap["xy"] == np.array(cell)).all(axis=1))[0].tolist())
        assert sim.grid_entities_at(0, cell) == want
    assert sim.grid_num_entities(0) == int(snap["alive"].sum())


def test_large_population_invariants():
    """5 M people at the reference density (1.25 per cell): conservation and cadence properties."""
    n, G = 5_000_000, 2000
    cfg = SpatialConfig(population=n, grid_size=G)
    sim, sc, stats = build(cfg)
    sim.run(40)
    ts = sim.get_aggregate_time_series(stats)
    v = np.array([astuple(x) for x in ts.values()], dtype=np.int64)
    assert len(ts) == 40 and ts.time() == list(range(1, 41))
    assert (v[:, :4].sum(axis=1) == v[:, 5]).all() and (v[:, 4] + v[:, 5] == n).all()
    assert (np.diff(v[:, 5]) <= 0).all() and (np.diff(v[:, 3]) >= 0).all()      # nobody is resurrected / un-recovers
    xy = sim.read_component(sc.position)
    assert xy.min() >= 0 and xy.max() < G
    q = sim.read_component(sc.quarantined)
    assert q.max() <= 14 and (q > 0).mean() > 0.5
