"""Developer micro-benchmarks (device-timed through incerto_sim_last_run_stats)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import incerto_b200 as ib
from incerto_b200 import _ffi as F

SEED = 0x1CE27002025
HBM = 6534.1


def forest(W, H, steps, series=1):
    b = ib.SimulationBuilder(seed=SEED)
    pos = b.component("pos", F.I16, lanes=2)
    cell = b.component("cell", F.U8)
    b.add_spatial_grid_2d(pos, cell, bounds=((0, 0), (W - 1, H - 1)))
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_FOREST_GRID, F.SpawnForestParams(pos.handle, cell.handle, W, H, 0.7, 3, 3)))
    b.add_systems(ib.FireSpread(pos, cell, 0.6, 3), ib.BurnProgression(pos, cell), ib.Regrowth(pos, cell, 0.001))
    if series:
        b.record_aggregate_time_series(ib.FireStatsOf(cell), series)
    sim = b.build()
    sim.run(20)
    t0 = time.perf_counter()
    sim.run(steps)
    wall = time.perf_counter() - t0
    ms, launches = sim.last_run_stats()
    upd = W * H * steps
    print(f"forest {W}x{H} steps={steps}: device {ms:.3f} ms ({ms / steps * 1e3:.2f} us/step), wall {wall * 1e3:.2f} ms, "
          f"{upd / ms / 1e6:.1f} G cell-updates/s, {2 * upd / ms / 1e6:.1f} GB/s algorithmic = {2 * upd / ms / 1e6 / HBM:.3f} of measured HBM; launches {launches}")
    fs = sim.sample_aggregate(ib.FireStatsOf(cell))
    print("   stats", fs.healthy_count, fs.burning_count, fs.burned_count, fs.empty_count)
    return sim


def coin(n, steps, fused):
    b = ib.SimulationBuilder(seed=SEED)
    heads = b.component("h", F.U32)
    tosses = b.component("t", F.U32)
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_COIN_TOSSERS, F.SpawnCoinTossersParams(heads.handle, tosses.handle, n)))
    b.add_systems(ib.CoinToss(heads, tosses, 0.5))
    sim = b.build()
    sim.run(3)
    if fused:
        sim.run(steps)
        ms, launches = sim.last_run_stats()
    else:
        ms = 0.0
        launches = 0
        for _ in range(steps):
            sim.run(1)
            m, l = sim.last_run_stats()
            ms += m
            launches += l
    upd = n * steps
    print(f"coin n={n} steps={steps} fused={fused}: device {ms:.3f} ms, {upd / ms / 1e6:.1f} G updates/s, "
          f"{16 * upd / ms / 1e6:.1f} GB/s algorithmic = {16 * upd / ms / 1e6 / HBM:.3f} of measured HBM; launches {launches}")


if __name__ == "__main__":
    forest(4096, 4096, 200)
    forest(4096, 4096, 200, series=0)
    forest(16384, 16384, 20)
    forest(1024, 1024, 200)
    coin(1 << 28, 10, False)
    coin(1 << 28, 100, True)
