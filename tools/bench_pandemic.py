"""pandemic micro-benchmark: device time per kernel through the engine's profiling hooks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import incerto_b200 as ib
from incerto_b200 import _ffi as F

SEED = 0x1CE27002025
n, G = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, int(sys.argv[2]) if len(sys.argv) > 2 else 5000
b = ib.SimulationBuilder(seed=SEED)
person, infected = b.component("Person"), b.component("Infected")
position = b.component("Position", F.U16, lanes=2)
b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_PANDEMIC_PEOPLE, F.SpawnPandemicParams(person.handle, position.handle, infected.handle, n, G, 0.05)))
b.add_systems(ib.PeopleMove(person, position, infected, G), ib.PeopleInfectOthers(person, position, infected, G, 0.02),
              ib.PeopleInfectedMayDie(person, position, infected, G, 0.0005), ib.PeopleInfectedMayRecover(person, position, infected, G, 0.001))
sim = b.build()
sim.run(10)
sim.run(100)
ms, launches = sim.last_run_stats()
print(f"pandemic n={n} G={G}: 100 steps {ms:.3f} ms device, {n*100/ms/1e6:.1f} G person-updates/s, {10*n*100/ms/1e6:.1f} GB/s algorithmic = {10*n*100/ms/1e6/6534.1:.3f} of measured HBM; launches {launches}")
sim.set_profiling(True)
sim.run(50)
for k in sim.kernel_stats():
    if k["device_ms"] > 0:
        print(f"   {k['name']:22s} {k['device_ms']/50*1e3:9.2f} us/step")
print("   survivors, infected:", sim.count(ib.With(person)), sim.count(ib.With(infected)))
