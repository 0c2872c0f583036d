"""pandemic (examples/pandemic.rs) on the GPU vs the CPU oracle: bit-exact positions, infection flags, survivors."""
import numpy as np
import pytest

import incerto_b200 as ib
from incerto_b200 import _ffi as F
from conftest import SEED
from oracle.pandemic import PandemicOracle

pytestmark = pytest.mark.gpu


def build(n, G, replica=0, p_start=0.05, p_inf=0.02, p_rec=0.001, p_die=0.0005, systems=15, host=None):
    b = ib.SimulationBuilder(seed=SEED, replica=replica)
    person = b.component("Person")
    infected = b.component("Infected")
    position = b.component("Position", F.U16, lanes=2)
    if host is None:
        b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_PANDEMIC_PEOPLE, F.SpawnPandemicParams(
            person.handle, position.handle, infected.handle, n, G, p_start)))
    else:
        xy, inf = host
        b.add_entity_spawner(lambda s: s.spawn_batch(len(inf), {person: None, position: xy, infected: inf.astype(np.uint8)}))
    sysl = []
    if systems & 1:
        sysl.append(ib.PeopleMove(person, position, infected, G))
    if systems & 2:
        sysl.append(ib.PeopleInfectOthers(person, position, infected, G, p_inf))
    if systems & 4:
        sysl.append(ib.PeopleInfectedMayDie(person, position, infected, G, p_die))
    if systems & 8:
        sysl.append(ib.PeopleInfectedMayRecover(person, position, infected, G, p_rec))
    b.add_systems(*sysl)
    return b.build(), person, infected, position


def check_state(sim, o, person, infected, position):
    xy_o, alive_o, inf_o = o.state()
    xy, uids = sim.read_component(position, with_uids=True)
    inf = sim.read_component(infected, rows_of=position).astype(bool)
    assert np.array_equal(uids, np.nonzero(alive_o)[0])            # the same people survived
    assert np.array_equal(xy, xy_o[alive_o])
    assert np.array_equal(inf, inf_o[alive_o])
    # examples/pandemic.rs:74-78
    assert (sim.count(ib.With(person)), sim.count(ib.With(infected)),
            sim.count(ib.With(person), ib.Without(infected))) == o.counts()


def test_reference_config():
    """examples/pandemic.rs:18-33 as-is (1000 people on 50x50), first 3000 of its 100 000 steps."""
    sim, person, infected, position = build(1000, 50)
    o = PandemicOracle(SEED, 0, 1000, 50).spawn()
    check_state(sim, o, person, infected, position)                # spawn parity (:84-107)
    for chunk in (1, 1, 98, 900, 2000):
        sim.run(chunk)
        o.run(chunk)
        check_state(sim, o, person, infected, position)
    assert o.counts()[0] < 1000                                    # somebody died: despawn path exercised


@pytest.mark.parametrize("n,G,replica", [(1, 1, 0), (3, 2, 1), (4097, 17, 2), (20_000, 100, 3), (50_001, 300, 4)])
def test_parity_sizes(n, G, replica):
    sim, person, infected, position = build(n, G, replica=replica, p_inf=0.1, p_rec=0.01, p_die=0.005)
    o = PandemicOracle(SEED, replica, n, G, p_infect=0.1, p_recover=0.01, p_die=0.005, nested_loop=False).spawn()
    for chunk in (1, 19, 180):
        sim.run(chunk)
        o.run(chunk)
        check_state(sim, o, person, infected, position)


@pytest.mark.parametrize("systems", [1, 2, 3, 4, 8, 6, 10, 12, 14])
def test_system_subsets(systems):
    n, G = 5000, 20
    sim, person, infected, position = build(n, G, p_inf=0.05, p_rec=0.02, p_die=0.01, systems=systems)
    o = PandemicOracle(SEED, 0, n, G, p_infect=0.05, p_recover=0.02, p_die=0.01, systems=systems, nested_loop=False).spawn()
    sim.run(60)
    o.run(60)
    check_state(sim, o, person, infected, position)


def test_one_level_fate_draws():
    """Probabilities above 1/256 use the single-level draw of the contract."""
    n, G = 3000, 10
    sim, person, infected, position = build(n, G, p_inf=0.3, p_rec=0.2, p_die=0.1)
    o = PandemicOracle(SEED, 0, n, G, p_infect=0.3, p_recover=0.2, p_die=0.1, nested_loop=False).spawn()
    sim.run(40)
    o.run(40)
    check_state(sim, o, person, infected, position)


def test_host_spawn_crowded_cell_and_walls():
    """Everybody in the corner cells: many infected per cell, saturating moves at the walls (:131,:141)."""
    n, G = 600, 6
    rng = np.random.default_rng(1)
    xy = np.where(rng.random((n, 1)) < 0.5, 0, G - 1).astype(np.uint16) * np.ones((1, 2), np.uint16)
    inf = rng.random(n) < 0.3
    sim, person, infected, position = build(n, G, p_inf=0.02, host=(xy, inf))
    o = PandemicOracle(SEED, 0, n, G, nested_loop=True).spawn_explicit(xy, inf)
    for _ in range(30):
        sim.run(1)
        o.run(1)
        check_state(sim, o, person, infected, position)


def test_more_than_2_pow_24_people():
    """Populations of 2^24 - 1 rows and more use 32-bit list links (no step stamp in the heads, which are cleared before
    every step) instead of the 24-bit links of smaller ones: same results as the oracle, person by person."""
    n, G = (1 << 24) + 12345, 6500
    sim, person, infected, position = build(n, G)
    o = PandemicOracle(SEED, 0, n, G, nested_loop=False).spawn()
    sim.run(3)
    o.run(3)
    check_state(sim, o, person, infected, position)


def test_step_tag_wraparound():
    """The per-cell list heads are stamped with step & 0xff: run across several wrap-arounds."""
    n, G = 2000, 8
    sim, person, infected, position = build(n, G, p_inf=0.01, p_rec=0.02, p_die=0.0)
    o = PandemicOracle(SEED, 0, n, G, p_infect=0.01, p_recover=0.02, p_die=0.0, nested_loop=False).spawn()
    sim.run(700)
    o.run(700)
    check_state(sim, o, person, infected, position)


def test_replica_reset_and_large_population_invariants():
    n, G = 2_000_000, 2236   # density 0.4 like the reference (1000 / 50^2)
    sim, person, infected, position = build(n, G)
    sim.run(20)
    c1 = (sim.count(ib.With(person)), sim.count(ib.With(infected)))
    xy = sim.read_component(position)
    assert xy.max() < G and sim.count(ib.With(person)) + 0 <= n
    assert sim.count(ib.With(person)) == sim.count(ib.With(infected)) + sim.count(ib.With(person), ib.Without(infected))
    sim.reset(0)
    sim.run(20)
    assert (sim.count(ib.With(person)), sim.count(ib.With(infected))) == c1
    assert np.array_equal(sim.read_component(position), xy)
    # a window of the population against the oracle is not possible (people interact), so compare a small twin
    sim.reset(5)
    sim.run(5)
    assert (sim.count(ib.With(person)), sim.count(ib.With(infected))) != c1
