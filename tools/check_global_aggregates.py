"""Run under torchrun: incerto_sim_sample_aggregate_global (aggregates over the entities of all ranks; Median /
Percentile as a distributed radix select over NCCL) against numpy on the concatenated data.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_global_aggregates.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import incerto_b200 as ib
from incerto_b200 import _ffi as F
from incerto_b200 import dist as D

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def data(r, dtype, n):
    rng = np.random.default_rng(1000 + 17 * r)
    if np.issubdtype(dtype, np.floating):
        return (rng.standard_normal(n) * 50 + 3 * r).astype(dtype)
    info = np.iinfo(dtype)
    lo, hi = max(info.min, -100000), min(info.max, 100000)
    return rng.integers(lo, hi, n, endpoint=True).astype(dtype)


ok = True
for dt_code, np_dt in ((F.F64, np.float64), (F.I32, np.int32), (F.U16, np.uint16), (F.F32, np.float32), (F.I64, np.int64)):
    sizes = [5000 + 1234 * r for r in range(world)]          # ragged: every rank holds a different number of entities
    sizes[-1] = 7                                            # and one holds almost nothing
    mine = data(rank, np_dt, sizes[rank])
    b = ib.SimulationBuilder(seed=1, device=local, replica=rank)
    v = b.component("v", dt_code)
    ga = b.component("GroupA")
    in_a = (np.arange(sizes[rank]) % 3 == 0).astype(np.uint8)
    b.add_entity_spawner(lambda s_, mine=mine, in_a=in_a: s_.spawn_batch(len(mine), {v: mine, ga: in_a}))
    sim = b.build()
    sim.attach_comm(D.share_unique_id(dist, ib.comm_unique_id, rank), rank, world)
    everything = np.concatenate([data(r, np_dt, sizes[r]) for r in range(world)])
    group = np.concatenate([data(r, np_dt, sizes[r])[np.arange(sizes[r]) % 3 == 0] for r in range(world)])
    for name, sel, filt in (("all", everything, ()), ("With<GroupA>", group, (ib.With(ga),))):
        srt = np.sort(sel)
        n = len(srt)
        want = {"count": n, "min": srt[0], "max": srt[-1], "median": srt[n // 2],
                "p25": srt[int(np.floor(25 / 100 * n))], "p90": srt[int(np.floor(90 / 100 * n))]}
        got = {"count": sim.sample_aggregate_global(ib.Count(v).filtered(*filt)),
               "min": sim.sample_aggregate_global(ib.Minimum(v).filtered(*filt)),
               "max": sim.sample_aggregate_global(ib.Maximum(v).filtered(*filt)),
               "median": sim.sample_aggregate_global(ib.Median(v).filtered(*filt)),
               "p25": sim.sample_aggregate_global(ib.Percentile(v, 25).filtered(*filt)),
               "p90": sim.sample_aggregate_global(ib.Percentile(v, 90).filtered(*filt))}
        good = all(np_dt(got[k]) == np_dt(want[k]) if k != "count" else int(got[k]) == want[k] for k in want)
        if np.issubdtype(np_dt, np.floating):
            mean = sim.sample_aggregate_global(ib.Mean(v).filtered(*filt))
            good = good and abs(mean - float(np.mean(sel.astype(np.float64)))) <= 1e-6 * (1 + abs(float(np.mean(sel))))
        ok = ok and good
        print(f"rank {rank}/{world} {np_dt.__name__:8s} {name:13s} global aggregates {'ok' if good else 'MISMATCH ' + str(got) + ' != ' + str(want)}", flush=True)
    # the local aggregate is still the local one
    assert int(sim.sample_aggregate(ib.Count(v))) == sizes[rank]
    sim.destroy()

# coin_toss: the global odds over all ranks' entities (every rank another replica)
b = ib.SimulationBuilder(seed=0x1CE27002025, device=local, replica=rank)
heads, tosses = b.component("h", F.U32), b.component("t", F.U32)
n = 1000 + 100 * rank
b.add_entity_spawner(lambda s_: s_.spawn_batch(n, {heads: np.zeros(n, np.uint32), tosses: np.zeros(n, np.uint32)}))
b.add_systems(ib.CoinToss(heads, tosses, 0.5))
sim = b.build()
sim.attach_comm(D.share_unique_id(dist, ib.comm_unique_id, rank), rank, world)
sim.run(40)
local_odds = sim.sample_aggregate(ib.CoinOdds(heads, tosses))
t = torch.tensor([local_odds * n, n], dtype=torch.float64, device="cuda")
dist.all_reduce(t)
want = float(t[0] / t[1])
got = sim.sample_aggregate_global(ib.CoinOdds(heads, tosses))
good = abs(got - want) < 1e-12
ok = ok and good
print(f"rank {rank}/{world} coin odds over all replicas: {got:.6f} ({'ok' if good else 'MISMATCH ' + str(want)})", flush=True)
sim.destroy()
flag = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
