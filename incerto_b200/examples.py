"""The reference's example programs (`examples/*.rs`, their `main()` builder chains) restated through the façade,
parameterised by population so that bench.py, the tools and the tests build exactly the same simulations.

Each function returns `(sim, handles)` where `handles` names the components / aggregates the example reads back.
All populations are spawned by the built-in device spawners (the examples' spawn functions under the Philox contract,
DESIGN.md §2); pass `replica` to run another independent Monte Carlo replica of the same configuration.
"""
from __future__ import annotations

from types import SimpleNamespace

from . import _ffi as F
from . import simulation as ib

SEED = 0x1CE27002025


def coin_toss(n: int, replica: int = 0, device: int | None = None, series: int = 1, seed: int = SEED):
    """examples/coin_toss.rs:20-56 (NUM_COIN_TOSSERS = n, ODDS_HEADS = 0.5)."""
    b = ib.SimulationBuilder(seed=seed, replica=replica, device=device)
    heads, tosses = b.component("num_heads", F.U32), b.component("num_tosses", F.U32)
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_COIN_TOSSERS, F.SpawnCoinTossersParams(heads.handle, tosses.handle, n)))
    b.add_systems(ib.CoinToss(heads, tosses, 0.5))
    odds = ib.CoinOdds(heads, tosses)
    if series:
        b.record_aggregate_time_series(odds, series)      # one launch per sampled step instead of a fused run(n)
    return b.build(), SimpleNamespace(heads=heads, tosses=tosses, odds=odds, n=n)


def forest_fire(W: int, H: int, replica: int = 0, device: int | None = None, series: int = 1, seed: int = SEED,
                partition: tuple[int, int] | None = None, density: float = 0.7, p_spread: float = 0.6, burn: int = 3,
                p_regrow: float = 0.001, fires: int = 3):
    """examples/forest_fire.rs:147-170; `partition` = (rank, world) splits the grid by columns of x."""
    b = ib.SimulationBuilder(seed=seed, replica=replica, device=device)
    pos = b.component("GridPosition2D", F.I16, lanes=2)
    cell = b.component("ForestCell.state", F.U8)
    if partition is not None and partition[1] > 1:
        b.set_row_partition(*partition)
    b.add_spatial_grid_2d(pos, cell, bounds=((0, 0), (W - 1, H - 1)))
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_FOREST_GRID, F.SpawnForestParams(pos.handle, cell.handle, W, H, density, burn, fires)))
    b.add_systems(ib.FireSpread(pos, cell, p_spread, burn), ib.BurnProgression(pos, cell), ib.Regrowth(pos, cell, p_regrow))
    stats = ib.FireStatsOf(cell)
    if series:
        b.record_aggregate_time_series(stats, series)
    return b.build(), SimpleNamespace(pos=pos, cell=cell, stats=stats, n=W * H)


def pandemic(n: int, G: int, replica: int = 0, device: int | None = None, seed: int = SEED):
    """examples/pandemic.rs:56-81 (INITIAL_POPULATION = n, GRID_SIZE = G)."""
    b = ib.SimulationBuilder(seed=seed, replica=replica, device=device)
    person, infected = b.component("Person"), b.component("Infected")
    position = b.component("Position", F.U16, lanes=2)
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_PANDEMIC_PEOPLE, F.SpawnPandemicParams(person.handle, position.handle, infected.handle, n, G, 0.05)))
    b.add_systems(ib.PeopleMove(person, position, infected, G), ib.PeopleInfectOthers(person, position, infected, G, 0.02),
                  ib.PeopleInfectedMayDie(person, position, infected, G, 0.0005),
                  ib.PeopleInfectedMayRecover(person, position, infected, G, 0.001))
    return b.build(), SimpleNamespace(person=person, infected=infected, position=position, n=n)


def pandemic_spatial(n: int, G: int, replica: int = 0, device: int | None = None, series: int = 1, seed: int = SEED):
    """examples/pandemic_spatial.rs:160-200 with the constants of :28-55."""
    b = ib.SimulationBuilder(seed=seed, replica=replica, device=device)
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
    if series:
        b.record_aggregate_time_series(stats, series)
    return b.build(), SimpleNamespace(sc=sc, stats=stats, n=n)


def traders(nt: int, ns: int = 100, replica: int = 0, device: int | None = None, seed: int = SEED):
    """examples/traders.rs:76-100 (NUM_TRADERS = nt, NUM_STOCKS = ns)."""
    b = ib.SimulationBuilder(seed=seed, replica=replica, device=device)
    tc = ib.TraderComponents(b.component("TraderId", F.U64), b.component("cash", F.F64), b.component("portfolio", F.U8),
                             b.component("TraderNetWorth", F.F64), b.component("StockId", F.U64),
                             b.component("StockPrice", F.F64), b.component("StockDelisted"))
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_TRADERS, F.SpawnTradersParams(
        tc.trader_id.handle, tc.cash.handle, tc.portfolio.handle, tc.net_worth.handle, nt, 100.0, ns)))
    b.add_systems(ib.TradersMayBuyShares(tc), ib.TradersMaySellShares(tc), ib.TradersCalculateNetWorth(tc))
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_STOCKS, F.SpawnStocksParams(
        tc.stock_id.handle, tc.stock_price.handle, tc.stock_delisted.handle, ns, 5.0)))
    b.add_systems(ib.StocksPriceChange(tc))
    return b.build(), SimpleNamespace(tc=tc, n=nt)


def air_traffic_3d(n: int, size: int, height: int = 10, replica: int = 0, device: int | None = None, series: int = 1,
                   seed: int = SEED, partition: tuple[int, int] | None = None):
    """examples/air_traffic_3d.rs:201-220 (AIRSPACE_SIZE = size, AIRSPACE_HEIGHT = height); `partition` = (rank, world)
    splits the ONE airspace by x over the ranks (aircraft migrate between them)."""
    b = ib.SimulationBuilder(seed=seed, replica=replica, device=device)
    position = b.component("GridPosition3D", F.I16, lanes=3)
    target = b.component("Aircraft.target_altitude", F.I16)
    if partition is not None and partition[1] > 1:
        b.set_row_partition(*partition)
    b.add_spatial_grid_3d(position, target, bounds=((0, 0, 0), (size, size, height)))
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_AIRCRAFT, F.SpawnAircraftParams(position.handle, target.handle, n, size, height)))
    b.add_systems(ib.MoveAircraft(position, target, size, height), ib.CheckConflicts(position, target, size, height))
    conflicts = ib.AirConflicts(target)
    if series:
        b.record_aggregate_time_series(conflicts, series)
    return b.build(), SimpleNamespace(position=position, target=target, conflicts=conflicts, n=n)
