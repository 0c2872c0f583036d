import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from incerto_b200 import examples as ex
sim, h = ex.forest_fire(64, 512)
sim.run(400)          # the fire has burnt out; the three igniter entities are still there
for _ in range(40):
    sim.run(1)
