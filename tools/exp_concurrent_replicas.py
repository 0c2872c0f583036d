"""Do independent replicas running concurrently on one GPU (one engine handle and stream each) raise the throughput of
the latency-bound workloads?   python tools/exp_concurrent_replicas.py [pandemic|air|traders] """
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import incerto_b200 as ib
from incerto_b200 import _ffi as F

SEED = 0x1CE27002025
workload = sys.argv[1] if len(sys.argv) > 1 else "pandemic"
steps = 100


def make(replica):
    b = ib.SimulationBuilder(seed=SEED, replica=replica)
    if workload == "pandemic":
        n, G = 10_000_000, 5000
        person, infected = b.component("Person"), b.component("Infected")
        position = b.component("Position", F.U16, lanes=2)
        b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_PANDEMIC_PEOPLE, F.SpawnPandemicParams(person.handle, position.handle, infected.handle, n, G, 0.05)))
        b.add_systems(ib.PeopleMove(person, position, infected, G), ib.PeopleInfectOthers(person, position, infected, G, 0.02),
                      ib.PeopleInfectedMayDie(person, position, infected, G, 0.0005), ib.PeopleInfectedMayRecover(person, position, infected, G, 0.001))
        return b.build(), n
    if workload == "traders":
        n, ns = 1_000_000, 100
        tc = ib.TraderComponents(b.component("TraderId", F.U64), b.component("cash", F.F64), b.component("portfolio", F.U8),
                                 b.component("TraderNetWorth", F.F64), b.component("StockId", F.U64),
                                 b.component("StockPrice", F.F64), b.component("StockDelisted"))
        b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_TRADERS, F.SpawnTradersParams(
            tc.trader_id.handle, tc.cash.handle, tc.portfolio.handle, tc.net_worth.handle, n, 100.0, ns)))
        b.add_systems(ib.TradersMayBuyShares(tc), ib.TradersMaySellShares(tc), ib.TradersCalculateNetWorth(tc))
        b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_STOCKS, F.SpawnStocksParams(
            tc.stock_id.handle, tc.stock_price.handle, tc.stock_delisted.handle, ns, 5.0)))
        b.add_systems(ib.StocksPriceChange(tc))
        return b.build(), n
    raise SystemExit("unknown workload")


for K in (1, 2, 3, 4):
    sims = [make(r) for r in range(K)]
    n = sims[0][1]
    for s, _ in sims:
        s.run(5)
    torch.cuda.synchronize()
    best = None
    for rep in range(3):
        t0 = time.perf_counter()
        for s, _ in sims:
            s.run_async(steps)
        for s, _ in sims:
            s.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    one = sims[0][0].last_run_stats()[0]
    print(f"{workload}: {K} concurrent replica(s): {best * 1e3:.2f} ms wall for {K} x {steps} steps -> {K * n * steps / best / 1e9:.1f} G updates/s "
          f"({best / (K * steps) * 1e6:.1f} us per replica-step; handle 0 alone saw {one / steps * 1e3:.1f} us/step on its stream)", flush=True)
    for s, _ in sims:
        s.destroy()
