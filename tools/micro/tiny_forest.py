"""Where does a forest step's fixed cost go?  Small and bench-sized grids stepped through captured graphs and cold,
with the overlay igniters / regrowth / spread removed in turn.   python tools/micro/tiny_forest.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import incerto_b200 as ib
from incerto_b200 import _ffi as F
from incerto_b200 import examples as ex


def build(W, H, fires=3, systems=7, series=1):
    b = ib.SimulationBuilder(seed=ex.SEED)
    pos = b.component("GridPosition2D", F.I16, lanes=2)
    cell = b.component("ForestCell.state", F.U8)
    b.add_spatial_grid_2d(pos, cell, bounds=((0, 0), (W - 1, H - 1)))
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_FOREST_GRID, F.SpawnForestParams(pos.handle, cell.handle, W, H, 0.7, 3, fires)))
    sysl = []
    if systems & 1: sysl.append(ib.FireSpread(pos, cell, 0.6, 3))
    if systems & 2: sysl.append(ib.BurnProgression(pos, cell))
    if systems & 4: sysl.append(ib.Regrowth(pos, cell, 0.001))
    b.add_systems(*sysl)
    if series:
        b.record_aggregate_time_series(ib.FireStatsOf(cell), series)
    return b.build()


for (W, H) in ((64, 512), (4096, 4096)):
    for name, kw in (("as bench", {}), ("no overlays", dict(fires=0)), ("no regrowth", dict(systems=3)), ("no overlays, no regrowth", dict(fires=0, systems=3)),
                     ("progress only", dict(fires=0, systems=2)), ("no series", dict(series=0))):
        sim = build(W, H, **kw)
        sim.run(64)
        sim.run(256)
        warm = sim.last_run_stats()[0] / 256 * 1e3
        sim.run_cold(5)
        sim.run_cold(20)
        cold = sim.last_run_stats()[0] / 20 * 1e3
        print(f"{W}x{H} {name:26s}: graph replay {warm:6.2f} us/step, cold {cold:6.2f} us/step", flush=True)
        sim.destroy()
