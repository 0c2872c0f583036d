"""Replica-sharded Monte Carlo over the GPUs of one box (BASELINE.json configs[2] and [4]): independent replicas of
pandemic (10 M agents) or traders (1 M traders x 100 stocks) are dealt round-robin to the ranks, run back to back on
each GPU by resetting a simulation handle under the next replica key (`incerto_sim_reset`) - `--concurrent K` (default
2) handles, each with its own stream, keep K replicas in flight per GPU - and the per-replica
aggregate statistics are reduced across ranks over NCCL (`incerto_sim_allreduce_*`) - the only collective on this
path.  Launch:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_replicas.py
                [pandemic|traders] [--replicas R] [--steps S] [--scale X]
Timing: CUDA events on the engine streams (start of a wave of K replicas to its latest end, `incerto_sim_span_ms`),
summed per rank, MAX over ranks."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import incerto_b200 as ib
from incerto_b200 import _ffi as F
from incerto_b200 import dist as D

SEED = 0x1CE27002025


def arg(name, default, cast=int):
    return cast(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def build(workload, scale, replica, device):
    b = ib.SimulationBuilder(seed=SEED, replica=replica, device=device)
    h = {}
    if workload == "pandemic":
        n, G = int(10_000_000 * scale), int(5000 * scale ** 0.5)
        person, infected = b.component("Person"), b.component("Infected")
        position = b.component("Position", F.U16, lanes=2)
        b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_PANDEMIC_PEOPLE, F.SpawnPandemicParams(person.handle, position.handle, infected.handle, n, G, 0.05)))
        b.add_systems(ib.PeopleMove(person, position, infected, G), ib.PeopleInfectOthers(person, position, infected, G, 0.02),
                      ib.PeopleInfectedMayDie(person, position, infected, G, 0.0005), ib.PeopleInfectedMayRecover(person, position, infected, G, 0.001))
        h.update(person=person, infected=infected)
        unit, bpu = "agent-updates/s", 10
    else:
        n, ns = int(1_000_000 * scale), 100
        tc = ib.TraderComponents(b.component("TraderId", F.U64), b.component("cash", F.F64), b.component("portfolio", F.U8),
                                 b.component("TraderNetWorth", F.F64), b.component("StockId", F.U64),
                                 b.component("StockPrice", F.F64), b.component("StockDelisted"))
        b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_TRADERS, F.SpawnTradersParams(
            tc.trader_id.handle, tc.cash.handle, tc.portfolio.handle, tc.net_worth.handle, n, 100.0, ns)))
        b.add_systems(ib.TradersMayBuyShares(tc), ib.TradersMaySellShares(tc), ib.TradersCalculateNetWorth(tc))
        b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_STOCKS, F.SpawnStocksParams(
            tc.stock_id.handle, tc.stock_price.handle, tc.stock_delisted.handle, ns, 5.0)))
        b.add_systems(ib.StocksPriceChange(tc))
        h.update(tc=tc)
        unit, bpu = "trader-updates/s", 124
    return b.build(), h, n, unit, bpu


def main():
    workload = next((a for a in sys.argv[1:] if a in ("pandemic", "traders")), "pandemic")
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scale = arg("--scale", 1.0, float)
    replicas = arg("--replicas", 8 * world)
    steps = arg("--steps", 100)
    # replicas in flight per GPU: one engine handle (= one stream) each.  The per-step kernels of these workloads are
    # bound by latency and instruction issue rather than by HBM, so two replicas interleaved on the SMs finish sooner
    # than one after the other (tools/exp_concurrent_replicas.py: +18 % for pandemic, +16 % for traders).
    conc = max(1, arg("--concurrent", 2))
    sims = [build(workload, scale, rank, local) for _ in range(conc)]
    n, unit, bpu = sims[0][2], sims[0][3], sims[0][4]
    if world > 1:
        # aggregate statistics are reduced through the first handle's communicator
        sims[0][0].attach_comm(D.share_unique_id(dist, ib.comm_unique_id, rank), rank, world)
    for sim, *_ in sims:
        sim.run(5)                              # warm-up on this rank's first replica
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = 0.0
    mine = list(range(rank, replicas, world))   # replica r -> rank r mod N
    u64 = np.zeros(2, np.uint64)                # pandemic: survivors, infected summed over replicas
    f64 = np.zeros(3, np.float64)               # traders: sum / min / max of the per-replica mean net worth
    f64[1], f64[2] = np.inf, -np.inf
    for w0 in range(0, len(mine), conc):
        wave = mine[w0:w0 + conc]
        for (sim, *_), r in zip(sims, wave):
            sim.reset(r)
        for (sim, *_), r in zip(sims, wave):
            sim.run_async(steps)
        for (sim, *_), r in zip(sims, wave):
            sim.synchronize()
        # device time of the wave: from the first handle's start to the latest end
        ms += max(sims[0][0].span_ms(sims[k][0]) for k in range(len(wave)))
        for (sim, h, *_), r in zip(sims, wave):
            if workload == "pandemic":
                u64 += np.array([sim.count(ib.With(h["person"])), sim.count(ib.With(h["infected"]))], np.uint64)
            else:
                m = sim.sample_aggregate(ib.Mean(h["tc"].net_worth))
                f64[0] += m
                f64[1], f64[2] = min(f64[1], m), max(f64[2], m)
    sim = sims[0][0]
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # cross-replica aggregate statistics over NCCL through the engine's communicator
        u64 = sim.allreduce_u64(u64, 0)
        f64[0:1] = sim.allreduce_f64(f64[0:1].copy(), 0)
        f64[1:2] = sim.allreduce_f64(f64[1:2].copy(), 1)
        f64[2:3] = sim.allreduce_f64(f64[2:3].copy(), 2)
        if workload == "traders":
            # a quantile over the traders of the replicas resident right now (one per rank): distributed radix select
            resident_median = sim.sample_aggregate_global(ib.Median(sims[0][1]["tc"].net_worth))
    if rank == 0:
        value = n * steps * replicas / (ms * 1e-3)
        out = {"workload": workload, "metric": "entity-updates/sec (device-timed)", "unit": unit, "value": value,
               "n_gpus": world, "replicas": replicas, "replicas_in_flight_per_gpu": conc, "entities_per_replica": n,
               "steps_per_replica": steps, "ms_max_over_ranks": ms, "scaling": "weak" if "--replicas" not in sys.argv else "strong",
               "algorithmic_GBps_per_gpu": value * bpu / world / 1e9}
        if workload == "pandemic":
            out["survivors_all_replicas"], out["infected_all_replicas"] = int(u64[0]), int(u64[1])
        else:
            out["mean_net_worth"] = {"mean_over_replicas": f64[0] / replicas, "min": f64[1], "max": f64[2]}
            if world > 1:
                out["median_net_worth_over_the_resident_replicas"] = resident_median
        print(json.dumps(out), flush=True)
    for s_, *_ in sims:
        s_.destroy()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
