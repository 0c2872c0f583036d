"""How fast does the spawn-time row order of air_traffic_3d (rows sorted by (x, y) column) stop helping?  Device time of
each of the first steps after the spawn, then of steps 50.. (the regime bench.py measures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from incerto_b200 import examples as ex
sim, h = ex.air_traffic_3d(5_000_000, 11547)
sim.run(1)   # first step: also builds the first bitmap
t = []
for k in range(24):
    sim.run(1)
    t.append(sim.last_run_stats()[0] * 1e3)
print("steps 2..25 (us):", " ".join(f"{v:.0f}" for v in t))
sim.run(30)
sim.run(20)
print("steps 56..75: %.1f us/step" % (sim.last_run_stats()[0] / 20 * 1e3))
os.environ["X"] = "1"
