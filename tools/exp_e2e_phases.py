"""Where does the e2e job of bench.py spend its time?  (spawn from host arrays -> run(500) -> read the series)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import incerto_b200 as ib
from incerto_b200 import _ffi as F

SEED = 0x1CE27002025
W = H = 4096
xs, ys = np.meshgrid(np.arange(W, dtype=np.int16), np.arange(H, dtype=np.int16), indexing="ij")
host_pos = np.concatenate([np.stack([xs.ravel(), ys.ravel()], 1), np.array([[7, 9], [2000, 2000], [4095, 0]], np.int16)])
rng = np.random.default_rng(1)
host_state = np.concatenate([np.where(rng.random(W * H) < 0.7, F.CELL_HEALTHY, F.CELL_EMPTY).astype(np.uint8), np.array([3, 3, 3], np.uint8)])
n = host_state.shape[0]
for it in range(3):
    t = [time.perf_counter()]
    b = ib.SimulationBuilder(seed=SEED)
    pos = b.component("GridPosition2D", F.I16, lanes=2)
    cell = b.component("ForestCell.state", F.U8)
    b.add_spatial_grid_2d(pos, cell, bounds=((0, 0), (W - 1, H - 1)))
    b.add_entity_spawner(lambda s: s.spawn_batch(n, {pos: host_pos, cell: host_state}))
    b.add_systems(ib.FireSpread(pos, cell, 0.6, 3), ib.BurnProgression(pos, cell), ib.Regrowth(pos, cell, 0.001))
    b.record_aggregate_time_series(ib.FireStatsOf(cell), 1)
    t.append(time.perf_counter())
    sim = b.build(borrow_spawn_buffers=True)
    t.append(time.perf_counter())
    sim.run(500)
    t.append(time.perf_counter())
    ts = sim.get_aggregate_time_series(ib.FireStatsOf(cell))
    fs = sim.sample_aggregate(ib.FireStatsOf(cell))
    t.append(time.perf_counter())
    sim.destroy()
    t.append(time.perf_counter())
    print("builder %.1f ms | build() %.1f ms | run(500) %.1f ms | series+aggregate %.1f ms | destroy %.1f ms | total %.1f ms" %
          tuple([(t[i + 1] - t[i]) * 1e3 for i in range(5)] + [(t[-1] - t[0]) * 1e3]))

