"""Per-workload device timings at the BASELINE.json sizes (one replica / one world per GPU).

For every workload: entity-updates/s over a timed run(steps) (CUDA events on the engine's stream, read through
incerto_sim_last_run_stats), the algorithmic GB/s that corresponds to (SURVEY §8d bytes per update) and its fraction
of the measured HBM peak, then a per-kernel breakdown from the engine's profiling hooks (events around each launch).
Usage: python tools/bench_workloads.py [coin forest pandemic spatial traders air] [--scale 1.0]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import incerto_b200 as ib
from incerto_b200 import _ffi as F

SEED = 0x1CE27002025
try:
    HBM = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6534.1
SCALE = 1.0


def report(name, sim, n, steps, bytes_per_update, warm=5, prof_steps=10):
    sim.run(warm)
    sim.run(steps)
    ms, launches = sim.last_run_stats()
    ups = n * steps / ms / 1e6
    print(f"{name}: n={n} steps={steps}: {ms / steps * 1e3:.1f} us/step, {ups:.2f} G updates/s, "
          f"{bytes_per_update} B/update -> {ups * bytes_per_update:.0f} GB/s algorithmic = "
          f"{ups * bytes_per_update / HBM:.3f} of measured HBM ({HBM:.0f} GB/s); {launches / steps:.1f} launches/step", flush=True)
    sim.set_profiling(True)
    base = {k["name"]: (k["device_ms"], k["launches"], k["algorithmic_bytes"]) for k in sim.kernel_stats()}
    sim.run(prof_steps)
    for k in sim.kernel_stats():
        b = base.get(k["name"], (0.0, 0, 0.0))
        dms, dl, db = k["device_ms"] - b[0], k["launches"] - b[1], k["algorithmic_bytes"] - b[2]
        if dl and dms > 0:
            print(f"   {k['name']:22s} {dms / dl * 1e3:9.2f} us/launch x {dl / prof_steps:.1f}/step   "
                  f"{db / dl / 1e6:9.1f} MB algorithmic/launch -> {db / dms / 1e6:8.0f} GB/s = {db / dms / 1e6 / HBM:.3f}")
    sim.set_profiling(False)


def coin():
    n = int((1 << 28) * SCALE)
    b = ib.SimulationBuilder(seed=SEED)
    heads, tosses = b.component("h", F.U32), b.component("t", F.U32)
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_COIN_TOSSERS, F.SpawnCoinTossersParams(heads.handle, tosses.handle, n)))
    b.add_systems(ib.CoinToss(heads, tosses, 0.5))
    b.record_aggregate_time_series(ib.CoinOdds(heads, tosses), 1)      # forces one launch per step
    report("coin_toss (series/1)", b.build(), n, 20, 16)


def forest():
    W = H = int(16384 * SCALE ** 0.5)
    b = ib.SimulationBuilder(seed=SEED)
    pos, cell = b.component("pos", F.I16, lanes=2), b.component("cell", F.U8)
    b.add_spatial_grid_2d(pos, cell, bounds=((0, 0), (W - 1, H - 1)))
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_FOREST_GRID, F.SpawnForestParams(pos.handle, cell.handle, W, H, 0.7, 3, 3)))
    b.add_systems(ib.FireSpread(pos, cell, 0.6, 3), ib.BurnProgression(pos, cell), ib.Regrowth(pos, cell, 0.001))
    b.record_aggregate_time_series(ib.FireStatsOf(cell), 1)
    report(f"forest_fire {W}x{H}", b.build(), W * H, 50, 2)


def pandemic():
    n, G = int(10_000_000 * SCALE), int(5000 * SCALE ** 0.5)
    b = ib.SimulationBuilder(seed=SEED)
    person, infected = b.component("Person"), b.component("Infected")
    position = b.component("Position", F.U16, lanes=2)
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_PANDEMIC_PEOPLE, F.SpawnPandemicParams(person.handle, position.handle, infected.handle, n, G, 0.05)))
    b.add_systems(ib.PeopleMove(person, position, infected, G), ib.PeopleInfectOthers(person, position, infected, G, 0.02),
                  ib.PeopleInfectedMayDie(person, position, infected, G, 0.0005),
                  ib.PeopleInfectedMayRecover(person, position, infected, G, 0.001))
    sim = b.build()
    report(f"pandemic G={G}", sim, n, 100, 10)
    print("   survivors, infected:", sim.count(ib.With(person)), sim.count(ib.With(infected)))


def spatial():
    n, G = int(50_000_000 * SCALE), int(6325 * SCALE ** 0.5)
    b = ib.SimulationBuilder(seed=SEED)
    position = b.component("GridPosition2D", F.I16, lanes=2)
    disease = b.component("Person.disease_state", F.U8)
    social = b.component("Person.social_distancing")
    quarantined = b.component("Quarantined.remaining_duration", F.U8)
    history = b.component("ContactHistory", F.U32)
    sc = ib.SpatialComponents(position, disease, social, quarantined, history, grid_size=G)
    b.add_spatial_grid_2d(position, disease, bounds=((0, 0), (G - 1, G - 1)))
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_SPATIAL_POPULATION, F.SpawnSpatialParams(
        position.handle, disease.handle, social.handle, quarantined.handle, history.handle, n, G, 0.7, 0.02, 5)))
    b.add_systems(ib.PeopleMoveWithSocialDistancing(sc), ib.DiseaseIncubationProgression(sc), ib.SpatialDiseaseTransmission(sc),
                  ib.DiseaseRecoveryAndDeath(sc), ib.UpdateContactHistory(sc), ib.ProcessContactTracing(sc),
                  ib.UpdateQuarantineStatus(sc))
    stats = ib.PandemicStatsOf(disease, n)
    b.record_aggregate_time_series(stats, 1)
    sim = b.build()
    report(f"pandemic_spatial G={G}", sim, n, 100, 104, warm=20)
    per_step = []
    for _ in range(32):
        sim.run(1)
        per_step.append(sim.last_run_stats()[0] * 1e3)
    print("   per-step us (quarantine waves, period 15):", " ".join(f"{v:.0f}" for v in per_step))
    ps = sim.sample_aggregate(stats)
    print("   stats:", ps.healthy_count, ps.exposed_count, ps.infectious_count, ps.recovered_count, ps.dead_count)


def traders():
    nt, ns = int(1_000_000 * SCALE), 100
    b = ib.SimulationBuilder(seed=SEED)
    tc = ib.TraderComponents(b.component("TraderId", F.U64), b.component("cash", F.F64), b.component("portfolio", F.U8),
                             b.component("TraderNetWorth", F.F64), b.component("StockId", F.U64),
                             b.component("StockPrice", F.F64), b.component("StockDelisted"))
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_TRADERS, F.SpawnTradersParams(
        tc.trader_id.handle, tc.cash.handle, tc.portfolio.handle, tc.net_worth.handle, nt, 100.0, ns)))
    b.add_systems(ib.TradersMayBuyShares(tc), ib.TradersMaySellShares(tc), ib.TradersCalculateNetWorth(tc))
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_STOCKS, F.SpawnStocksParams(
        tc.stock_id.handle, tc.stock_price.handle, tc.stock_delisted.handle, ns, 5.0)))
    b.add_systems(ib.StocksPriceChange(tc))
    sim = b.build()
    report("traders 100 stocks", sim, nt, 50, 124, warm=100)
    print("   mean / median / max net worth:", sim.sample_aggregate(ib.Mean(tc.net_worth)),
          sim.sample_aggregate(ib.Median(tc.net_worth)), sim.sample_aggregate(ib.Maximum(tc.net_worth)))


def air():
    n, size, height = int(5_000_000 * SCALE), int(11547 * SCALE ** 0.5), 10
    b = ib.SimulationBuilder(seed=SEED)
    position = b.component("GridPosition3D", F.I16, lanes=3)
    target = b.component("Aircraft.target_altitude", F.I16)
    b.add_spatial_grid_3d(position, target, bounds=((0, 0, 0), (size, size, height)))
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_AIRCRAFT, F.SpawnAircraftParams(position.handle, target.handle, n, size, height)))
    b.add_systems(ib.MoveAircraft(position, target, size, height), ib.CheckConflicts(position, target, size, height))
    b.record_aggregate_time_series(ib.AirConflicts(target), 1)
    sim = b.build()
    report(f"air_traffic_3d {size}x{size}x{height}", sim, n, 50, 18)
    print("   conflicts of the last step:", sim.sample_aggregate(ib.AirConflicts(target)))


def cpu():
    """The CPU oracle (structure-faithful C++ restatement of the Bevy systems, 1 thread - Bevy runs each system's loop on
    one core) on bounded samples of the same workloads: the baseline the GPU numbers are reported beside."""
    import time
    from oracle.air import AirOracle
    from oracle.forest import ForestOracle, coin_toss_run
    from oracle.pandemic import PandemicOracle
    from oracle.spatial import SpatialConfig, SpatialOracle
    from oracle.traders import TradersOracle

    def timed(name, n, steps, make, note=""):
        o = make()
        t0 = time.perf_counter()
        o.run(steps) if hasattr(o, "run") else None
        dt = time.perf_counter() - t0
        print(f"cpu {name}: n={n} steps={steps}: {n * steps / dt / 1e6:.2f} M updates/s on 1 core ({dt:.1f} s){note}", flush=True)

    t0 = time.perf_counter()
    coin_toss_run(SEED, 0, 1_000_000, 0.5, 1, 20)
    dt = time.perf_counter() - t0
    print(f"cpu coin_toss: n=1000000 steps=20: {20e6 / dt / 1e6:.2f} M updates/s on 1 core ({dt:.1f} s)", flush=True)
    timed("forest_fire 1024x1024", 1024 * 1024, 10, lambda: ForestOracle(SEED, 0, 1024, 1024).spawn())
    timed("pandemic G=707", 200_000, 20, lambda: PandemicOracle(SEED, 0, 200_000, 707, nested_loop=False).spawn(),
          " [per-cell buckets; the reference's O(infected x healthy) double loop is far slower]")
    timed("pandemic_spatial G=400", 200_000, 30, lambda: SpatialOracle(SEED, 0, SpatialConfig(population=200_000, grid_size=400)).spawn())
    timed("traders 100 stocks", 20_000, 50, lambda: TradersOracle(SEED, 0, 20_000, 100))
    timed("air_traffic_3d 2309x2309x10", 200_000, 20, lambda: AirOracle(SEED, 0, 200_000, 2309, 10).spawn())


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if "--scale" in sys.argv:
        SCALE = float(sys.argv[sys.argv.index("--scale") + 1])
        args = [a for a in args if a != sys.argv[sys.argv.index("--scale") + 1]]
    todo = args or ["coin", "forest", "pandemic", "spatial", "traders", "air"]
    for w in todo:
        globals()[w]()
