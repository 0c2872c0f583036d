"""Experiment: how much do the spatial workloads gain when rows are stored in cell order?  (host-side sort before spawn)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import incerto_b200 as ib
from incerto_b200 import _ffi as F

SEED = 0x1CE27002025


def spatial(n, G, sort):
    rng = np.random.default_rng(5)
    xy = rng.integers(0, G, size=(n, 2)).astype(np.int16)
    if sort:
        order = np.argsort(xy[:, 0].astype(np.int64) * G + xy[:, 1], kind="stable")
        xy = xy[order]
    d = np.where(rng.random(n) < 0.02, 5, 0).astype(np.uint8)
    sd = (rng.random(n) < 0.7).astype(np.uint8)
    b = ib.SimulationBuilder(seed=SEED)
    position = b.component("GridPosition2D", F.I16, lanes=2)
    disease = b.component("Person.disease_state", F.U8)
    social = b.component("Person.social_distancing")
    quarantined = b.component("Quarantined.remaining_duration", F.U8)
    history = b.component("ContactHistory", F.U32)
    sc = ib.SpatialComponents(position, disease, social, quarantined, history, grid_size=G)
    b.add_spatial_grid_2d(position, disease, bounds=((0, 0), (G - 1, G - 1)))
    b.add_entity_spawner(lambda s: s.spawn_batch(n, {position: xy, disease: d, social: sd, quarantined: None, history: None}))
    b.add_systems(ib.PeopleMoveWithSocialDistancing(sc), ib.DiseaseIncubationProgression(sc), ib.SpatialDiseaseTransmission(sc),
                  ib.DiseaseRecoveryAndDeath(sc), ib.UpdateContactHistory(sc), ib.ProcessContactTracing(sc),
                  ib.UpdateQuarantineStatus(sc))
    sim = b.build()
    sim.run(20)
    sim.set_profiling(True)
    sim.run(45)
    print(f"spatial n={n} G={G} sorted={sort}")
    for k in sim.kernel_stats():
        if k["device_ms"] > 0:
            print(f"   {k['name']:22s} {k['device_ms'] / k['launches'] * 1e3:9.2f} us/launch")


def pandemic(n, G, sort):
    rng = np.random.default_rng(5)
    xy = rng.integers(0, G, size=(n, 2)).astype(np.uint16)
    if sort:
        order = np.argsort(xy[:, 0].astype(np.int64) * G + xy[:, 1], kind="stable")
        xy = xy[order]
    inf = (rng.random(n) < 0.05).astype(np.uint8)
    b = ib.SimulationBuilder(seed=SEED)
    person, infected = b.component("Person"), b.component("Infected")
    position = b.component("Position", F.U16, lanes=2)
    b.add_entity_spawner(lambda s: s.spawn_batch(n, {person: None, position: xy, infected: inf}))
    b.add_systems(ib.PeopleMove(person, position, infected, G), ib.PeopleInfectOthers(person, position, infected, G, 0.02),
                  ib.PeopleInfectedMayDie(person, position, infected, G, 0.0005),
                  ib.PeopleInfectedMayRecover(person, position, infected, G, 0.001))
    sim = b.build()
    sim.run(10)
    sim.set_profiling(True)
    sim.run(50)
    print(f"pandemic n={n} G={G} sorted={sort}")
    for k in sim.kernel_stats():
        if k["device_ms"] > 0:
            print(f"   {k['name']:22s} {k['device_ms'] / k['launches'] * 1e3:9.2f} us/launch")


if __name__ == "__main__":
    import sys
    n, G = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (20_000_000, 4000)
    for s in (False, True):
        spatial(n, G, s)
    for s in (False, True):
        pandemic(10_000_000, 5000, s)
