"""Cold-L2 step time of forest_fire for a list of WxH grids on one GPU (incerto_sim_run_cold)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import incerto_b200 as ib
from incerto_b200 import _ffi as F

SEED = 0x1CE27002025
HBM = 6534.1
for spec in sys.argv[1:] or ["4096x4096", "2048x16384", "16384x16384", "1024x1024"]:
    W, H = (int(v) for v in spec.split("x"))
    b = ib.SimulationBuilder(seed=SEED)
    pos = b.component("pos", F.I16, lanes=2)
    cell = b.component("cell", F.U8)
    b.add_spatial_grid_2d(pos, cell, bounds=((0, 0), (W - 1, H - 1)))
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_FOREST_GRID, F.SpawnForestParams(pos.handle, cell.handle, W, H, 0.7, 3, 3)))
    b.add_systems(ib.FireSpread(pos, cell, 0.6, 3), ib.BurnProgression(pos, cell), ib.Regrowth(pos, cell, 0.001))
    b.record_aggregate_time_series(ib.FireStatsOf(cell), 1)
    sim = b.build()
    sim.run_cold(10)
    steps = 100
    sim.run_cold(steps)
    ms, _ = sim.last_run_stats()
    sim.run(64)
    sim.run(256)
    wm, _ = sim.last_run_stats()
    print(f"forest {W}x{H}: cold {ms / steps * 1e3:.2f} us/step = {2 * W * H * steps / ms / 1e6:.0f} GB/s ({2 * W * H * steps / ms / 1e6 / HBM:.3f}); "
          f"L2-warm run(256) {wm / 256 * 1e3:.2f} us/step", flush=True)
    sim.destroy()
