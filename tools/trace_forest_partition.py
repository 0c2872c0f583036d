"""Per-warp, per-stage timeline of one cold step of the ROW-SPLIT forest (one slab per GPU, fused halo transport), to set
beside the same slab alone on one GPU.  Traced build first (python tools/trace_forest.py --build), then on the box:
  INCERTO_B200_LIB=incerto_b200/lib/libincerto_b200.trace.so python -m torch.distributed.run --nproc-per-node 2 \
      --master-addr 127.0.0.1 --master-port 29533 tools/trace_forest_partition.py [slab_w x H]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import incerto_b200 as ib
from incerto_b200 import _ffi as F

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
spec = next((a for a in sys.argv[1:] if "x" in a), "2048x16384")
SW, H = (int(v) for v in spec.split("x"))
W = SW * world
b = ib.SimulationBuilder(seed=0x1CE27002025, device=local)
pos = b.component("pos", F.I16, lanes=2)
cell = b.component("cell", F.U8)
b.add_spatial_grid_2d(pos, cell, bounds=((0, 0), (W - 1, H - 1)))
b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_FOREST_GRID, F.SpawnForestParams(pos.handle, cell.handle, W, H, 0.7, 3, 3)))
b.add_systems(ib.FireSpread(pos, cell, 0.6, 3), ib.BurnProgression(pos, cell), ib.Regrowth(pos, cell, 0.001))
b.record_aggregate_time_series(ib.FireStatsOf(cell), 1)
if world > 1:
    b.set_row_partition(rank, world)
sim = b.build()
if world > 1:
    box = [ib.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    sim.attach_comm(box[0], rank, world)
sim.run_cold(20)
lib = F.load()
lib.incerto_debug_forest_trace.argtypes = [C.c_void_p, C.c_uint64]
for rep in range(3):
    sim.run_cold(1)
    ms, _ = sim.last_run_stats()
    ctas = ((H + 511) // 512) * ((SW + 27) // 28)
    nwarps = min(ctas * 4, 1 << 16)
    buf = np.zeros(8 * nwarps, np.uint64)
    assert lib.incerto_debug_forest_trace(buf.ctypes.data, 8 * nwarps) == 0
    t = buf.reshape(-1, 8).astype(np.int64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    start, loaded, end = t[:, 0] - t0, t[:, 1] - t0, t[:, 2] - t0
    stored, lot, events, overlays = t[:, 4] - t0, t[:, 5] - t0, t[:, 6] - t0, t[:, 7] - t0
    q = [0, 10, 50, 90, 99, 100]
    pct = lambda a: [round(float(np.percentile(a, p)) / 1e3, 2) for p in q]
    late = np.argsort(end)[-4:]
    if rep == 2:
        print(f"rank {rank}/{world} slab {SW}x{H}: event {ms * 1e3:.2f} us, span {end.max() / 1e3:.2f} us; start {pct(start)} loaded {pct(loaded)} end {pct(end)} "
              f"after-loads {pct(end - loaded)}; latest warps (index: start/loaded/stored/events/end): "
              + "; ".join(f"{w}: {start[w] / 1e3:.2f}/{loaded[w] / 1e3:.2f}/{stored[w] / 1e3:.2f}/{events[w] / 1e3:.2f}/{end[w] / 1e3:.2f}" for w in late), flush=True)
sim.destroy()
if world > 1:
    dist.destroy_process_group()
