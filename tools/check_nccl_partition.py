"""Run under torchrun: row-partitioned forest over NCCL (real halo exchange) vs the CPU oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import incerto_b200 as ib
from incerto_b200 import _ffi as F
from incerto_b200 import dist as D
from oracle.forest import ForestOracle

SEED = 0x1CE27002025
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, steps = 64 * world, 96, 80
if "--narrow" in sys.argv:   # slabs narrower than one CTA column group: both boundaries live in the same CTAs
    W, H, steps = 9 * world, 700, 60
if "--two-groups" in sys.argv:   # exactly two column groups per slab
    W, H, steps = 27 * world, 300, 60
RESET = "--reset" in sys.argv
CUT = "--cut" in sys.argv   # host-spawned grid with overlay entities on both sides of every slab boundary
b = ib.SimulationBuilder(seed=SEED, device=local)
pos = b.component("GridPosition2D", F.I16, lanes=2)
cell = b.component("ForestCell.state", F.U8)
b.set_row_partition(rank, world)
b.add_spatial_grid_2d(pos, cell, bounds=((0, 0), (W - 1, H - 1)))
if CUT:
    rng = np.random.default_rng(3)
    xs, ys = np.meshgrid(np.arange(W), np.arange(H), indexing="ij")
    xy = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.int32)
    st = np.where(rng.random(W * H) < 0.9, 0x40, 0x80).astype(np.uint8)
    ov_xy, ov_st = [[0, 0], [W - 1, H - 1]], [3, 3]
    for r in range(1, world):
        cut = ib.partition_range(W, r, world)[0]
        ov_xy += [[cut - 1, 10], [cut, 20], [cut, 20], [cut - 1, 30], [cut, 31], [cut - 1, 60], [cut, 61]]
        ov_st += [3, 3, 2, 0x40, 3, 0x40, 0x40]
    ov_xy, ov_st = np.array(ov_xy, np.int32), np.array(ov_st, np.uint8)
    all_xy, all_st = np.concatenate([xy, ov_xy]), np.concatenate([st, ov_st])
    b.add_entity_spawner(lambda s_: s_.spawn_batch(len(all_st), {pos: all_xy.astype(np.int16), cell: all_st}))
else:
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_FOREST_GRID, F.SpawnForestParams(pos.handle, cell.handle, W, H, 0.85, 3, 6)))
b.add_systems(ib.FireSpread(pos, cell, 0.6, 3), ib.BurnProgression(pos, cell), ib.Regrowth(pos, cell, 0.004))
b.record_aggregate_time_series(ib.FireStatsOf(cell), 1)
sim = b.build()
sim.attach_comm(D.share_unique_id(dist, ib.comm_unique_id, rank), rank, world)
sim.run(steps)
replica = 0
if RESET:
    # reset under another replica key and run again (exercises the quiesce + barrier of the peer transport)
    replica = 3
    sim.reset(replica)
    sim.run(steps)
_names = {k["name"] for k in sim.kernel_stats() if k["launches"]}
transport = "fused" if "forest_step_fused_halo" in _names else "peer" if "forest_halo_push_peer" in _names else "nccl"
if CUT:
    o = ForestOracle(SEED, replica, W, H, p_regrow=0.004, fires=0).spawn_explicit(all_xy, all_st).run(steps)
else:
    o = ForestOracle(SEED, replica, W, H, density=0.85, p_regrow=0.004, fires=6).spawn().run(steps)
want = o.cells()[0]
xb, xe = ib.partition_range(W, rank, world)
full = sim.read_component(cell)
got = full[:(xe - xb) * H]
ok = np.array_equal(got, want[xb * H:xe * H])
if CUT:
    # the overlay entities this rank owns (the slab is followed by the full overlay array)
    for k, (x, _) in enumerate(ov_xy):
        if xb <= x < xe:
            ok = ok and full[(xe - xb) * H + k] == want[W * H + k]
fs = sim.sample_aggregate(ib.FireStatsOf(cell))          # all-reduced over the communicator
stats_ok = (fs.healthy_count, fs.burning_count, fs.burned_count, fs.empty_count, fs.total_cells) == tuple(int(v) for v in o.stats())
ts = sim.get_aggregate_time_series(ib.FireStatsOf(cell))
series_ok = np.array_equal(np.array([(v.healthy_count, v.burning_count, v.burned_count, v.empty_count, v.total_cells) for v in ts.values()], dtype=np.uint64), o.series())
print(f"rank {rank}/{world}: halo transport {transport}, slab parity {ok}, all-reduced stats {stats_ok}, series {series_ok}, burned {int(o.stats()[2])}", flush=True)
flag = torch.tensor([int(ok and stats_ok and series_ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
