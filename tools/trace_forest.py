"""Per-warp timeline of one cold forest step (developer tool).  Build the traced variant first (here, no GPU needed):
    python tools/trace_forest.py --build
then on the GPU box:  INCERTO_B200_LIB=incerto_b200/lib/libincerto_b200.trace.so python tools/trace_forest.py [WxH]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if "--build" in sys.argv:
    from incerto_b200 import _build
    print(_build.build(extra_flags=["-DINCERTO_FOREST_TRACE"], variant="trace"))
    sys.exit(0)
import numpy as np
import incerto_b200 as ib
from incerto_b200 import _ffi as F

spec = next((a for a in sys.argv[1:] if "x" in a), "4096x4096")
W, H = (int(v) for v in spec.split("x"))
warm = "--warm" in sys.argv
fires = int(sys.argv[sys.argv.index("--fires") + 1]) if "--fires" in sys.argv else 3
b = ib.SimulationBuilder(seed=0x1CE27002025)
pos = b.component("pos", F.I16, lanes=2)
cell = b.component("cell", F.U8)
b.add_spatial_grid_2d(pos, cell, bounds=((0, 0), (W - 1, H - 1)))
b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_FOREST_GRID, F.SpawnForestParams(pos.handle, cell.handle, W, H, 0.7, 3, fires)))
b.add_systems(ib.FireSpread(pos, cell, 0.6, 3), ib.BurnProgression(pos, cell), ib.Regrowth(pos, cell, 0.001))
b.record_aggregate_time_series(ib.FireStatsOf(cell), 1)
sim = b.build()
sim.run_cold(20)
for rep in range(3):
    if warm:
        sim.run(1)
    else:
        sim.run_cold(1)
    ms, _ = sim.last_run_stats()
    lib = F.load()
    tile_rows, warps_per_cta, cols_per_thread = 512, 4, 7     # kTX * 16, kTY, kColsPerThread of kernels_forest.cu
    ctas = ((H + tile_rows - 1) // tile_rows) * ((W + warps_per_cta * cols_per_thread - 1) // (warps_per_cta * cols_per_thread))
    nwarps = min(ctas * warps_per_cta, 1 << 16)
    buf = np.zeros(8 * nwarps, np.uint64)
    lib.incerto_debug_forest_trace.argtypes = [C.c_void_p, C.c_uint64]
    assert lib.incerto_debug_forest_trace(buf.ctypes.data, 8 * nwarps) == 0
    t = buf.reshape(-1, 8).astype(np.int64)
    t0 = t[:, 0].min()
    start, loaded, end, sm = t[:, 0] - t0, t[:, 1] - t0, t[:, 2] - t0, t[:, 3] & 0xffffffff
    fire = t[:, 3] >> 32
    stored, lot, events, overlays = t[:, 4] - t0, t[:, 5] - t0, t[:, 6] - t0, t[:, 7] - t0
    q = [0, 1, 10, 25, 50, 75, 90, 99, 100]
    print(f"== {W}x{H} {'warm' if warm else 'cold'}: event time {ms * 1e3:.2f} us; warps {nwarps}; span first start -> last end {end.max() / 1e3:.2f} us")
    print("   percentiles      ", q)
    print("   warp start  (us) ", [round(float(np.percentile(start, p)) / 1e3, 2) for p in q])
    print("   1st tile in (us) ", [round(float(np.percentile(loaded, p)) / 1e3, 2) for p in q])
    print("   warp end    (us) ", [round(float(np.percentile(end, p)) / 1e3, 2) for p in q])
    print("   load wait   (us) ", [round(float(np.percentile(loaded - start, p)) / 1e3, 2) for p in q])
    print("   after loads (us) ", [round(float(np.percentile(end - loaded, p)) / 1e3, 2) for p in q])
    for name, a, b_ in (("compute+store", loaded, stored), ("lottery tail", stored, lot), ("event loop", lot, events), ("overlays", events, overlays),
                        ("stats+exit", overlays, end)):
        print(f"   {name:14s}(us) ", [round(float(np.percentile(b_ - a, p)) / 1e3, 2) for p in q], " fire warps:",
              [round(float(np.percentile((b_ - a)[fire > 0], p)) / 1e3, 2) for p in (50, 100)] if (fire > 0).any() else "-")
    late = np.argsort(end)[-6:]
    for w in late:
        print(f"      late warp {w} (cta {w // warps_per_cta}, sm {sm[w]}, fire {fire[w]}): start {start[w] / 1e3:.2f} loaded {loaded[w] / 1e3:.2f} stored {stored[w] / 1e3:.2f} "
              f"lottery {lot[w] / 1e3:.2f} events {events[w] / 1e3:.2f} overlays {overlays[w] / 1e3:.2f} end {end[w] / 1e3:.2f}")
    if nwarps <= 64:
        for w in range(nwarps):
            print(f"      warp {w:3d} (cta {w // warps_per_cta}, sm {sm[w]}, fire {fire[w]}): start {start[w] / 1e3:6.2f}  loaded {loaded[w] / 1e3:6.2f}  stored {stored[w] / 1e3:6.2f}  "
                  f"lottery {lot[w] / 1e3:6.2f}  events {events[w] / 1e3:6.2f}  overlays {overlays[w] / 1e3:6.2f}  end {end[w] / 1e3:6.2f} us")
    # per SM: first start, last end, warps
    sms = np.unique(sm)
    first = np.array([start[sm == s].min() for s in sms]) / 1e3
    last = np.array([end[sm == s].max() for s in sms]) / 1e3
    cnt = np.array([(sm == s).sum() for s in sms])
    print(f"   SMs {len(sms)}: first start min/median/max {first.min():.2f}/{np.median(first):.2f}/{first.max():.2f} us; "
          f"last end min/median/max {last.min():.2f}/{np.median(last):.2f}/{last.max():.2f} us; warps per SM min/max {cnt.min()}/{cnt.max()}")
