"""The small seeded cases behind tests/golden/workloads.npz: one per workload, each runnable on the oracle
(`run_oracle`) and on the CUDA engine (`run_engine`) with identical outputs."""
import numpy as np

SEED = 0x1CE27002025
CASES = ["coin_toss", "forest_fire", "forest_fire_lottery", "pandemic", "pandemic_two_level", "pandemic_spatial", "traders", "air_traffic_3d", "normal_draws"]


def run_oracle(name):
    if name == "coin_toss":
        from oracle.forest import coin_toss_run
        h, t, agg = coin_toss_run(SEED, 3, 500, 0.5, 1, 200)
        return dict(heads=h.astype(np.uint32), tosses=t.astype(np.uint32), odds=np.array([agg]))
    if name == "forest_fire":
        from oracle.forest import ForestOracle
        o = ForestOracle(SEED, 1, 48, 40, density=0.7, p_regrow=0.004, fires=3).spawn().run(60)
        return dict(cells=o.cells()[0], series=o.series())
    if name == "forest_fire_lottery":
        # p_regrow <= 1/256: level 1 of the regrowth draw is the block lottery (256 uids per call; H = 100 puts the
        # blocks across column ends)
        from oracle.forest import ForestOracle
        o = ForestOracle(SEED, 7, 36, 100, density=0.6, p_regrow=0.002, fires=3).spawn().run(90)
        return dict(cells=o.cells()[0], series=o.series())
    if name in ("pandemic", "pandemic_two_level"):
        from oracle.pandemic import PandemicOracle
        if name == "pandemic":
            o = PandemicOracle(SEED, 2, 1500, 20, p_infect=0.05, p_recover=0.01, p_die=0.005, nested_loop=False).spawn().run(120)
        else:  # both fate probabilities <= 1/256: the two-level fate draws of the contract
            o = PandemicOracle(SEED, 8, 3000, 30, p_infect=0.05, p_recover=0.003, p_die=0.002, nested_loop=False).spawn().run(150)
        xy, alive, inf = o.state()
        return dict(xy=xy, alive=alive, infected=inf, counts=np.array(o.counts(), np.uint64))
    if name == "pandemic_spatial":
        from oracle.spatial import SpatialConfig, SpatialOracle
        o = SpatialOracle(SEED, 4, SpatialConfig(population=1200, grid_size=30, chance_start_infected=0.05, chance_die=0.01,
                                                 zone_radius=4)).spawn().run(70)
        st = o.state()
        return dict(xy=st["xy"].astype(np.int16), alive=st["alive"], disease=st["disease"], quarantined=st["quarantined"].astype(np.uint8),
                    history=st["history"], series=o.series())
    if name == "traders":
        from oracle.traders import TradersOracle
        o = TradersOracle(SEED, 5, 64, 40, p_buy=0.03, p_sell=0.02, std_dev=0.4).run(150)
        st = o.state()
        return dict(cash=st["cash"], net_worth=st["net_worth"], shares=st["shares"], price=st["price"], delisted=st["delisted"],
                    series7=o.series(7))
    if name == "air_traffic_3d":
        from oracle.air import AirOracle
        o = AirOracle(SEED, 6, 600, 12, 5).spawn().run(80)
        xyz, tgt = o.state()
        return dict(xyz=xyz.astype(np.int16), target=tgt.astype(np.int16), conflicts=o.series())
    if name == "normal_draws":
        from oracle.traders import det_normal
        rng = np.random.default_rng(11)
        words = rng.integers(0, 2**32, size=(256, 4), dtype=np.uint64).astype(np.uint32)
        words[0] = 0
        words[1] = 0xFFFFFFFF
        return dict(words=words, z=np.array([det_normal(w) for w in words]))
    raise KeyError(name)


def run_engine(name):
    """The same cases through the C ABI on the GPU (no oracle involved)."""
    import incerto_b200 as ib
    from incerto_b200 import _ffi as F
    if name == "coin_toss":
        b = ib.SimulationBuilder(seed=SEED, replica=3)
        heads, tosses = b.component("num_heads", F.U32), b.component("num_tosses", F.U32)
        b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_COIN_TOSSERS, F.SpawnCoinTossersParams(heads.handle, tosses.handle, 500)))
        b.add_systems(ib.CoinToss(heads, tosses, 0.5))
        sim = b.build()
        sim.run(200)
        return dict(heads=sim.read_component(heads), tosses=sim.read_component(tosses),
                    odds=np.array([sim.sample_aggregate(ib.CoinOdds(heads, tosses))]))
    if name in ("forest_fire", "forest_fire_lottery"):
        W, H, replica, density, p_regrow, steps = (48, 40, 1, 0.7, 0.004, 60) if name == "forest_fire" else (36, 100, 7, 0.6, 0.002, 90)
        b = ib.SimulationBuilder(seed=SEED, replica=replica)
        pos, cell = b.component("GridPosition2D", F.I16, lanes=2), b.component("ForestCell.state", F.U8)
        b.add_spatial_grid_2d(pos, cell, bounds=((0, 0), (W - 1, H - 1)))
        b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_FOREST_GRID, F.SpawnForestParams(pos.handle, cell.handle, W, H, density, 3, 3)))
        b.add_systems(ib.FireSpread(pos, cell, 0.6, 3), ib.BurnProgression(pos, cell), ib.Regrowth(pos, cell, p_regrow))
        b.record_aggregate_time_series(ib.FireStatsOf(cell), 1)
        sim = b.build()
        sim.run(steps)
        ts = sim.get_aggregate_time_series(ib.FireStatsOf(cell))
        ser = np.array([(v.healthy_count, v.burning_count, v.burned_count, v.empty_count, v.total_cells) for v in ts.values()], np.uint64)
        return dict(cells=sim.read_component(cell), series=ser)
    if name in ("pandemic", "pandemic_two_level"):
        n, G, replica, p_rec, p_die, steps = (1500, 20, 2, 0.01, 0.005, 120) if name == "pandemic" else (3000, 30, 8, 0.003, 0.002, 150)
        b = ib.SimulationBuilder(seed=SEED, replica=replica)
        person, infected = b.component("Person"), b.component("Infected")
        position = b.component("Position", F.U16, lanes=2)
        b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_PANDEMIC_PEOPLE, F.SpawnPandemicParams(person.handle, position.handle, infected.handle, n, G, 0.05)))
        b.add_systems(ib.PeopleMove(person, position, infected, G), ib.PeopleInfectOthers(person, position, infected, G, 0.05),
                      ib.PeopleInfectedMayDie(person, position, infected, G, p_die), ib.PeopleInfectedMayRecover(person, position, infected, G, p_rec))
        sim = b.build()
        sim.run(steps)
        xy_live, uids = sim.read_component(position, with_uids=True)
        inf_live = sim.read_component(infected, rows_of=position)
        alive = np.zeros(n, bool)
        alive[uids.astype(np.int64)] = True
        return dict(xy_live=xy_live, alive=alive, infected_live=inf_live.astype(bool),
                    counts=np.array([sim.count(ib.With(person)), sim.count(ib.With(infected)),
                                     sim.count(ib.With(person), ib.Without(infected))], np.uint64))
    if name == "pandemic_spatial":
        n, G = 1200, 30
        b = ib.SimulationBuilder(seed=SEED, replica=4)
        position = b.component("GridPosition2D", F.I16, lanes=2)
        disease = b.component("Person.disease_state", F.U8)
        social = b.component("Person.social_distancing")
        quarantined = b.component("Quarantined.remaining_duration", F.U8)
        history = b.component("ContactHistory", F.U32)
        sc = ib.SpatialComponents(position, disease, social, quarantined, history, grid_size=G, chance_die=0.01, quarantine_radius=4)
        b.add_spatial_grid_2d(position, disease, bounds=((0, 0), (G - 1, G - 1)))
        b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_SPATIAL_POPULATION, F.SpawnSpatialParams(
            position.handle, disease.handle, social.handle, quarantined.handle, history.handle, n, G, 0.7, 0.05, 5)))
        b.add_systems(ib.PeopleMoveWithSocialDistancing(sc), ib.DiseaseIncubationProgression(sc), ib.SpatialDiseaseTransmission(sc),
                      ib.DiseaseRecoveryAndDeath(sc), ib.UpdateContactHistory(sc), ib.ProcessContactTracing(sc), ib.UpdateQuarantineStatus(sc))
        stats = ib.PandemicStatsOf(disease, n)
        b.record_aggregate_time_series(stats, 1)
        sim = b.build()
        sim.run(70)
        xy_live, uids = sim.read_component(position, with_uids=True)
        alive = np.zeros(n, bool)
        alive[uids.astype(np.int64)] = True
        ts = sim.get_aggregate_time_series(stats)
        ser = np.array([(v.healthy_count, v.exposed_count, v.infectious_count, v.recovered_count, v.dead_count, v.total_population)
                        for v in ts.values()], np.uint64)
        return dict(xy_live=xy_live, alive=alive, disease_live=sim.read_component(disease),
                    quarantined_live=sim.read_component(quarantined), series=ser)
    if name == "traders":
        nt, ns = 64, 40
        b = ib.SimulationBuilder(seed=SEED, replica=5)
        tc = ib.TraderComponents(b.component("TraderId", F.U64), b.component("cash", F.F64), b.component("portfolio", F.U8),
                                 b.component("TraderNetWorth", F.F64), b.component("StockId", F.U64),
                                 b.component("StockPrice", F.F64), b.component("StockDelisted"))
        b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_TRADERS, F.SpawnTradersParams(
            tc.trader_id.handle, tc.cash.handle, tc.portfolio.handle, tc.net_worth.handle, nt, 100.0, ns)))
        b.add_systems(ib.TradersMayBuyShares(tc, 0.03), ib.TradersMaySellShares(tc, 0.02), ib.TradersCalculateNetWorth(tc))
        b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_STOCKS, F.SpawnStocksParams(
            tc.stock_id.handle, tc.stock_price.handle, tc.stock_delisted.handle, ns, 5.0)))
        b.add_systems(ib.StocksPriceChange(tc, 0.001, 0.4))
        b.record_time_series(tc.net_worth, tc.trader_id, 1)
        sim = b.build()
        sim.run(150)
        return dict(cash=sim.read_component(tc.cash), net_worth=sim.read_component(tc.net_worth),
                    shares=sim.read_portfolio(tc.portfolio, tc.cash, ns), price=sim.read_component(tc.stock_price),
                    delisted=sim.read_component(tc.stock_delisted, rows_of=tc.stock_price).astype(bool),
                    series7=np.array(sim.get_time_series(tc.net_worth, tc.trader_id, 7).values()))
    if name == "air_traffic_3d":
        n, size, height = 600, 12, 5
        b = ib.SimulationBuilder(seed=SEED, replica=6)
        position = b.component("GridPosition3D", F.I16, lanes=3)
        target = b.component("Aircraft.target_altitude", F.I16)
        b.add_spatial_grid_3d(position, target, bounds=((0, 0, 0), (size, size, height)))
        b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_AIRCRAFT, F.SpawnAircraftParams(position.handle, target.handle, n, size, height)))
        b.add_systems(ib.MoveAircraft(position, target, size, height), ib.CheckConflicts(position, target, size, height))
        b.record_aggregate_time_series(ib.AirConflicts(target), 1)
        sim = b.build()
        sim.run(80)
        ts = sim.get_aggregate_time_series(ib.AirConflicts(target))
        return dict(xyz=sim.read_component(position), target=sim.read_component(target), conflicts=np.array(ts.values(), np.uint64))
    raise KeyError(name)


def compare_engine_to_golden(name, got, gold):
    """gold: dict of the arrays stored under `name/` in workloads.npz."""
    def eq(a, b):
        a, b = np.asarray(a), np.asarray(b)
        if a.dtype.kind == "f" or b.dtype.kind == "f":
            return np.array_equal(a.astype(np.float64).view(np.uint64), b.astype(np.float64).view(np.uint64))
        return a.shape == b.shape and np.array_equal(a.astype(np.int64), b.astype(np.int64))
    if name in ("pandemic", "pandemic_two_level", "pandemic_spatial"):
        alive = gold["alive"].astype(bool)
        assert eq(got["alive"], alive), "survivors differ"
        assert eq(got["xy_live"], gold["xy"][alive])
        if name != "pandemic_spatial":
            assert eq(got["infected_live"], gold["infected"][alive]) and eq(got["counts"], gold["counts"])
        else:
            assert eq(got["disease_live"], gold["disease"][alive]) and eq(got["quarantined_live"], gold["quarantined"][alive])
            assert eq(got["series"], gold["series"])
        return
    if name == "coin_toss":
        assert eq(got["heads"], gold["heads"]) and eq(got["tosses"], gold["tosses"])
        assert abs(got["odds"][0] - gold["odds"][0]) < 1e-12
        return
    for k in gold:
        assert eq(got[k], gold[k]), f"{name}/{k} differs from the golden vector"
