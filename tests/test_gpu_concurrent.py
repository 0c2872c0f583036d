"""Independent replicas in flight on one GPU: one engine handle (one stream) each, driven with run_async.  The results
must be those of the same replicas run one after the other, and incerto_sim_span_ms must cover both runs."""
import numpy as np
import pytest

import incerto_b200 as ib
from incerto_b200 import _ffi as F
from conftest import SEED

pytestmark = pytest.mark.gpu


def build_pandemic(replica, n=200_000, G=700):
    b = ib.SimulationBuilder(seed=SEED, replica=replica)
    person, infected = b.component("Person"), b.component("Infected")
    position = b.component("Position", F.U16, lanes=2)
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_PANDEMIC_PEOPLE, F.SpawnPandemicParams(person.handle, position.handle, infected.handle, n, G, 0.05)))
    b.add_systems(ib.PeopleMove(person, position, infected, G), ib.PeopleInfectOthers(person, position, infected, G, 0.2),
                  ib.PeopleInfectedMayDie(person, position, infected, G, 0.005), ib.PeopleInfectedMayRecover(person, position, infected, G, 0.01))
    return b.build(), person, infected, position


def test_two_replicas_in_flight_equal_two_sequential_runs():
    steps = 40
    seq = []
    for r in (0, 1):
        sim, person, infected, position = build_pandemic(r)
        sim.run(steps)
        seq.append((sim.count(ib.With(person)), sim.count(ib.With(infected)), sim.read_component(position).copy()))
        sim.destroy()
    a = build_pandemic(0)
    b = build_pandemic(1)
    a[0].run_async(steps)
    b[0].run_async(steps)
    a[0].synchronize()
    b[0].synchronize()
    for (sim, person, infected, position), want in zip((a, b), seq):
        assert (sim.count(ib.With(person)), sim.count(ib.With(infected))) == want[:2]
        assert np.array_equal(sim.read_component(position), want[2])
    span = a[0].span_ms(b[0])
    ms_a, ms_b = a[0].last_run_stats()[0], b[0].last_run_stats()[0]
    assert span > 0 and span >= max(ms_a, ms_b) * 0.99 and span <= (ms_a + ms_b) * 1.5 + 1.0
    # a handle that is still running cannot be measured
    a[0].run_async(1)
    with pytest.raises(ib.EngineError):
        a[0].span_ms(b[0])
    a[0].synchronize()
