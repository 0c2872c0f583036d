"""Python mirror of incerto's `SimulationBuilder` / `Simulation` over the C ABI.

Same method names, argument meaning and error behaviour as the reference
(src/simulation_builder.rs:39-306, src/simulation.rs:22-259, src/error.rs), so the
parity tests read like the reference's own tests.  Everything here is host-side
plumbing: all simulation work happens inside libincerto_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Callable, Iterable, Sequence

import numpy as np

from . import _ffi as F

_NP = {F.U8: np.uint8, F.U16: np.uint16, F.U32: np.uint32, F.U64: np.uint64, F.I8: np.int8, F.I16: np.int16,
       F.I32: np.int32, F.I64: np.int64, F.F32: np.float32, F.F64: np.float64}


# ------------------------------------------------------------------ errors
class SimulationError(Exception):
    """src/error.rs:1-8"""


class SamplingError(SimulationError):
    """src/error.rs:12-43; `.kind` is the variant name."""
    _NAMES = {
        F.ERR_COMPONENT_DOES_NOT_EXIST: "ComponentDoesNotExist",
        F.ERR_SINGLE_NO_ENTITIES: "SingleNoEntities",
        F.ERR_SINGLE_MULTIPLE_ENTITIES: "SingleMultipleEntities",
        F.ERR_AGGREGATE_NO_ENTITIES: "AggregateNoEntities",
        F.ERR_TIME_SERIES_NOT_RECORDED: "TimeSeriesNotRecorded",
        F.ERR_ENTITY_IDENTIFIER_NOT_FOUND: "EntityIdentifierNotFound",
        F.ERR_ENTITY_IDENTIFIER_NOT_UNIQUE: "EntityIdentifierNotUnique",
    }

    def __init__(self, code: int):
        self.code = code
        self.kind = self._NAMES[code]
        super().__init__(self.kind)


class BuilderError(SimulationError):
    """src/error.rs:47-52"""

    def __init__(self, code: int):
        self.code = code
        self.kind = "TimeSeriesRecordingConflict"
        super().__init__(self.kind)


class EngineError(SimulationError):
    """Engine status codes with no reference counterpart (the reference panics or cannot express them)."""

    def __init__(self, code: int, msg: str):
        self.code = code
        super().__init__(f"incerto_b200 status {code}: {msg}")


def _check(code: int) -> None:
    if code == F.OK:
        return
    if code in SamplingError._NAMES:
        raise SamplingError(code)
    if code == F.ERR_TIME_SERIES_RECORDING_CONFLICT:
        raise BuilderError(code)
    raise EngineError(code, F.last_error())


# -------------------------------------------------------------- components
@dataclass(frozen=True)
class Component:
    """Handle of a registered component (one column or one marker)."""
    handle: int
    name: str
    dtype: int
    lanes: int = 1

    @property
    def np_dtype(self):
        return _NP[self.dtype]


@dataclass(frozen=True)
class With:
    component: Component


@dataclass(frozen=True)
class Without:
    component: Component


def _filter_struct(filters: Iterable) -> tuple[F.Filter, list]:
    w = [f.component.handle for f in filters if isinstance(f, With)]
    wo = [f.component.handle for f in filters if isinstance(f, Without)]
    wa = (F.component_t * max(len(w), 1))(*w)
    woa = (F.component_t * max(len(wo), 1))(*wo)
    fs = F.Filter(C.cast(wa, C.POINTER(F.component_t)), len(w), C.cast(woa, C.POINTER(F.component_t)), len(wo))
    return fs, [wa, woa]


# -------------------------------------------------------------- aggregates
@dataclass(frozen=True)
class Aggregate:
    """`Out` of SampleAggregate<Out> (src/traits.rs:11-20) for component `component`."""
    component: Component
    kind: int
    percentile: int = 0
    component2: Component | None = None
    param: int = 0
    filters: tuple = ()

    def filtered(self, *filters) -> "Aggregate":
        return Aggregate(self.component, self.kind, self.percentile, self.component2, self.param, tuple(filters))

    def _desc(self):
        fs, keep = _filter_struct(self.filters)
        d = F.AggregateDesc(self.component.handle, self.component2.handle if self.component2 else -1, self.kind,
                            self.percentile, self.param, fs)
        return d, keep

    def _decode(self, buf: bytes):
        k = self.kind
        if k in (F.AGG_COUNT,):
            return int(np.frombuffer(buf, np.uint64, 1)[0])
        if k == F.AGG_SUM:
            dt = self.component.dtype
            t = np.float64 if dt in (F.F32, F.F64) else (np.int64 if F.I8 <= dt <= F.I64 else np.uint64)
            v = np.frombuffer(buf, t, 1)[0]
            return float(v) if t is np.float64 else int(v)
        if k in (F.AGG_MINIMUM, F.AGG_MAXIMUM, F.AGG_MEAN, F.AGG_MEDIAN, F.AGG_PERCENTILE):
            v = np.frombuffer(buf, self.component.np_dtype, 1)[0]
            return v.item()
        if k == F.AGG_COIN_ODDS:
            return float(np.frombuffer(buf, np.float64, 1)[0])
        if k == F.AGG_FIRE_STATS:
            return F.FireStats.from_buffer_copy(buf)
        if k == F.AGG_PANDEMIC_STATS:
            return F.PandemicStats.from_buffer_copy(buf)
        if k == F.AGG_AIR_CONFLICTS:
            return int(np.frombuffer(buf, np.uint64, 1)[0])
        raise ValueError(k)


def Sum(c): return Aggregate(c, F.AGG_SUM)
def Count(c): return Aggregate(c, F.AGG_COUNT)
def Minimum(c): return Aggregate(c, F.AGG_MINIMUM)          # src/util/aggregate.rs:73-88
def Maximum(c): return Aggregate(c, F.AGG_MAXIMUM)          # :90-105
def Mean(c): return Aggregate(c, F.AGG_MEAN)                # :157-192
def Median(c): return Aggregate(c, F.AGG_MEDIAN)            # :137-155
def Percentile(c, p: int): return Aggregate(c, F.AGG_PERCENTILE, percentile=p)  # :107-135
def CoinOdds(heads, tosses): return Aggregate(heads, F.AGG_COIN_ODDS, component2=tosses)   # coin_toss.rs:33-49
def FireStatsOf(cell): return Aggregate(cell, F.AGG_FIRE_STATS)                          # forest_fire.rs:104-134
def PandemicStatsOf(c, initial_population): return Aggregate(c, F.AGG_PANDEMIC_STATS, param=initial_population)
def AirConflicts(c): return Aggregate(c, F.AGG_AIR_CONFLICTS)   # `conflicts / 2` of the last step, air_traffic_3d.rs:130-161


# ------------------------------------------------------------------ systems
@dataclass
class System:
    kind: int
    params: C.Structure


def AddConst(target: Component, amount: int = 1, with_: Sequence[Component] = ()) -> System:
    p = F.AddConstParams()
    p.target = target.handle
    p.amount = amount
    for i, c in enumerate(with_):
        p.with_[i] = c.handle
    p.n_with = len(with_)
    return System(F.SYS_ADD_CONST, p)


def AddColumn(target: Component, source: Component) -> System:
    return System(F.SYS_ADD_COLUMN, F.AddColumnParams(target.handle, source.handle))


def CoinToss(num_heads: Component, num_tosses: Component, odds_heads: float) -> System:
    return System(F.SYS_COIN_TOSS, F.CoinTossParams(num_heads.handle, num_tosses.handle, odds_heads))


def FireSpread(position, cell, probability: float, burn_duration: int) -> System:
    return System(F.SYS_FOREST_SPREAD, F.ForestParams(position.handle, cell.handle, probability, burn_duration))


def BurnProgression(position, cell) -> System:
    return System(F.SYS_FOREST_BURN_PROGRESSION, F.ForestParams(position.handle, cell.handle, 0.0, 0))


def Regrowth(position, cell, probability: float) -> System:
    return System(F.SYS_FOREST_REGROWTH, F.ForestParams(position.handle, cell.handle, probability, 0))


# examples/pandemic.rs:115-223
def PeopleMove(person, position, infected, grid_size: int) -> System:
    return System(F.SYS_PANDEMIC_MOVE, F.PandemicParams(person.handle, position.handle, infected.handle, grid_size, 0.0))


def PeopleInfectOthers(person, position, infected, grid_size: int, chance: float) -> System:
    return System(F.SYS_PANDEMIC_INFECT, F.PandemicParams(person.handle, position.handle, infected.handle, grid_size, chance))


def PeopleInfectedMayDie(person, position, infected, grid_size: int, chance: float) -> System:
    return System(F.SYS_PANDEMIC_DIE, F.PandemicParams(person.handle, position.handle, infected.handle, grid_size, chance))


def PeopleInfectedMayRecover(person, position, infected, grid_size: int, chance: float) -> System:
    return System(F.SYS_PANDEMIC_RECOVER, F.PandemicParams(person.handle, position.handle, infected.handle, grid_size, chance))


# examples/pandemic_spatial.rs:316-625
@dataclass(frozen=True)
class SpatialComponents:
    """Person{disease_state, social_distancing}, Quarantined, ContactHistory, GridPosition2D handles + the constants
    of examples/pandemic_spatial.rs:28-58 shared by its seven systems."""
    position: Component
    disease_state: Component
    social_distancing: Component
    quarantined: Component
    contact_history: Component
    grid_size: int = 40
    infection_radius: int = 2
    chance_infect_distance_1: float = 0.15
    chance_infect_distance_2: float = 0.05
    chance_recover: float = 0.03
    chance_die: float = 0.001
    incubation_period: int = 5
    crowding_threshold: int = 8
    contact_quarantine_duration: int = 14
    quarantine_center: tuple | None = None     # (GRID_SIZE / 2, GRID_SIZE / 2)
    quarantine_radius: int = 8
    contact_history_clear_above: int = 20
    chance_attempt_move: float = 0.5

    def _params(self) -> "F.SpatialParams":
        cx, cy = self.quarantine_center or (self.grid_size // 2, self.grid_size // 2)
        p = F.SpatialParams()
        p.position, p.disease_state = self.position.handle, self.disease_state.handle
        p.social_distancing, p.quarantined = self.social_distancing.handle, self.quarantined.handle
        p.contact_history = self.contact_history.handle
        p.grid_size, p.infection_radius = self.grid_size, self.infection_radius
        p.chance_infect_distance_1, p.chance_infect_distance_2 = self.chance_infect_distance_1, self.chance_infect_distance_2
        p.chance_recover, p.chance_die = self.chance_recover, self.chance_die
        p.incubation_period, p.crowding_threshold = self.incubation_period, self.crowding_threshold
        p.contact_quarantine_duration = self.contact_quarantine_duration
        p.quarantine_center_x, p.quarantine_center_y, p.quarantine_radius = cx, cy, self.quarantine_radius
        p.contact_history_clear_above = self.contact_history_clear_above
        p.chance_attempt_move = self.chance_attempt_move
        return p


def PeopleMoveWithSocialDistancing(sc: SpatialComponents) -> System: return System(F.SYS_SPATIAL_MOVE, sc._params())
def DiseaseIncubationProgression(sc: SpatialComponents) -> System: return System(F.SYS_SPATIAL_INCUBATION, sc._params())
def SpatialDiseaseTransmission(sc: SpatialComponents) -> System: return System(F.SYS_SPATIAL_TRANSMISSION, sc._params())
def DiseaseRecoveryAndDeath(sc: SpatialComponents) -> System: return System(F.SYS_SPATIAL_RECOVERY_DEATH, sc._params())
def UpdateContactHistory(sc: SpatialComponents) -> System: return System(F.SYS_SPATIAL_CONTACT_HISTORY, sc._params())
def ProcessContactTracing(sc: SpatialComponents) -> System: return System(F.SYS_SPATIAL_CONTACT_TRACING, sc._params())
def UpdateQuarantineStatus(sc: SpatialComponents) -> System: return System(F.SYS_SPATIAL_QUARANTINE, sc._params())


# examples/traders.rs:140-263
@dataclass(frozen=True)
class TraderComponents:
    """The component handles the four traders systems share (Trader{cash, portfolio}, TraderNetWorth, StockPrice ...)."""
    trader_id: Component
    cash: Component
    portfolio: Component
    net_worth: Component
    stock_id: Component
    stock_price: Component
    stock_delisted: Component

    def _params(self, **kw) -> "F.TradersParams":
        p = F.TradersParams(self.trader_id.handle, self.cash.handle, self.portfolio.handle, self.net_worth.handle,
                            self.stock_id.handle, self.stock_price.handle, self.stock_delisted.handle)
        p.buy_shares_min, p.buy_shares_max = 1, 5
        for k, v in kw.items():
            setattr(p, k, v)
        return p


def TradersMayBuyShares(tc: TraderComponents, chance_buy: float = 0.01, shares=(1, 5)) -> System:
    return System(F.SYS_TRADERS_BUY, tc._params(chance_buy=chance_buy, buy_shares_min=shares[0], buy_shares_max=shares[1]))


def TradersMaySellShares(tc: TraderComponents, chance_sell: float = 0.005) -> System:
    return System(F.SYS_TRADERS_SELL, tc._params(chance_sell=chance_sell))


def TradersCalculateNetWorth(tc: TraderComponents) -> System:
    return System(F.SYS_TRADERS_NET_WORTH, tc._params())


def StocksPriceChange(tc: TraderComponents, mean: float = 0.001, std_dev: float = 0.2) -> System:
    return System(F.SYS_STOCKS_PRICE_CHANGE, tc._params(price_change_mean=mean, price_change_std_dev=std_dev))


# examples/air_traffic_3d.rs:93-162
def CheckConflicts(position, target_altitude, airspace_size: int, airspace_height: int) -> System:
    return System(F.SYS_AIR_CHECK_CONFLICTS, F.AirParams(position.handle, target_altitude.handle, airspace_size,
                                                         airspace_height, 0.0))


def MoveAircraft(position, target_altitude, airspace_size: int, airspace_height: int, chance_horizontal_move: float = 0.7) -> System:
    return System(F.SYS_AIR_MOVE, F.AirParams(position.handle, target_altitude.handle, airspace_size, airspace_height,
                                              chance_horizontal_move))


@dataclass
class DeviceSpawner:
    kind: int
    params: C.Structure


class Spawner:
    """src/spawner.rs:3-12.  `spawn` takes one bundle; `spawn_batch` many at once.

    Consecutive `spawn` calls whose bundles carry the same data components are
    shipped as one bulk call (markers become per-row presence bytes), which keeps
    the spawn order — and therefore the uids — of the reference loop."""

    def __init__(self):
        self._calls: list[tuple[int, dict]] = []   # (n, {Component: [ndarray chunks] | None})
        self._rows: list[dict] = []

    def spawn(self, bundle: dict) -> None:
        if self._rows:
            sig = {c for c in self._rows[0] if c.dtype != F.MARKER}
            if sig != {c for c in bundle if c.dtype != F.MARKER}:
                self._flush_rows()
        self._rows.append(dict(bundle))

    def _flush_rows(self) -> None:
        rows, self._rows = self._rows, []
        if not rows:
            return
        n = len(rows)
        comps = []
        for r in rows:
            for c in r:
                if c not in comps:
                    comps.append(c)
        cols = {}
        for c in comps:
            if c.dtype == F.MARKER:
                present = np.fromiter((1 if c in r else 0 for r in rows), dtype=np.uint8, count=n)
                cols[c] = None if present.all() else [present.reshape(n, 1)]
            else:
                cols[c] = [np.asarray([r[c] for r in rows], dtype=c.np_dtype).reshape(n, c.lanes)]
        self._calls.append((n, cols))

    def spawn_batch(self, n: int, columns: dict) -> None:
        self._flush_rows()
        cols = {}
        for comp, val in columns.items():
            if comp.dtype == F.MARKER:
                cols[comp] = None if val is None else [np.asarray(val, dtype=np.uint8).reshape(n, 1)]
            elif val is None:
                cols[comp] = None      # Default::default(): zeros / empty engine-managed container
            else:
                cols[comp] = [np.ascontiguousarray(np.asarray(val, dtype=comp.np_dtype)).reshape(n, comp.lanes)]
        self._calls.append((n, cols))


class SimulationBuilder:
    """src/simulation_builder.rs:25-306"""

    def __init__(self, seed: int | None = None, replica: int = 0, device: int | None = None):
        self._lib = F.load()
        h = C.c_void_p()
        _check(self._lib.incerto_builder_new(C.byref(h)))
        self._h = h
        self._spawners: list = []
        self._keep: list = []
        if seed is not None:
            _check(self._lib.incerto_builder_set_seed(self._h, seed, replica))
        if device is not None:
            _check(self._lib.incerto_builder_set_device(self._h, device))

    def __del__(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.incerto_builder_free(self._h)
            self._h = None

    # -- engine-side registration of component types (Rust: #[derive(Component)])
    def component(self, name: str, dtype: int = F.MARKER, lanes: int = 1) -> Component:
        out = F.component_t()
        _check(self._lib.incerto_builder_register_component(self._h, name.encode(), dtype, lanes, C.byref(out)))
        return Component(out.value, name, dtype, lanes)

    def set_row_partition(self, rank: int, world: int) -> "SimulationBuilder":
        _check(self._lib.incerto_builder_set_row_partition(self._h, rank, world))
        return self

    def add_systems(self, *systems: System) -> "SimulationBuilder":           # :66-70
        flat = []
        for s in systems:
            flat.extend(s if isinstance(s, (tuple, list)) else [s])
        arr = (F.SystemDesc * len(flat))()
        for i, s in enumerate(flat):
            arr[i].kind = s.kind
            arr[i].params = C.cast(C.pointer(s.params), C.c_void_p)
            arr[i].params_size = C.sizeof(s.params)
        _check(self._lib.incerto_builder_add_systems(self._h, arr, len(flat)))
        return self

    def add_spatial_grid(self, dim: int, position: Component, component: Component, bounds=None):   # :114-125
        out = C.c_int32()
        b = None
        if bounds is not None:
            lo, hi = bounds
            b = (C.c_int32 * (2 * dim))(*list(lo), *list(hi))
        _check(self._lib.incerto_builder_add_spatial_grid(self._h, dim, b, position.handle, component.handle,
                                                          C.byref(out)))
        self.last_grid = out.value
        return self

    def add_spatial_grid_2d(self, position, component, bounds=None):           # :130-133
        return self.add_spatial_grid(2, position, component, bounds)

    def add_spatial_grid_3d(self, position, component, bounds=None):           # :138-141
        return self.add_spatial_grid(3, position, component, bounds)

    def add_entity_spawner(self, entity_spawner) -> "SimulationBuilder":       # :154-158
        """`entity_spawner` is a callable taking a Spawner (as in the reference) or a DeviceSpawner."""
        self._spawners.append(entity_spawner)
        return self

    def record_aggregate_time_series(self, agg: Aggregate, sample_interval: int) -> "SimulationBuilder":   # :230-283
        d, keep = agg._desc()
        _check(self._lib.incerto_builder_record_aggregate_time_series(self._h, C.byref(d), sample_interval))
        return self

    def record_aggregate_time_series_filtered(self, agg: Aggregate, filters, sample_interval: int):
        return self.record_aggregate_time_series(agg.filtered(*filters), sample_interval)

    def record_time_series(self, component: Component, identifier: Component, sample_interval: int):       # :185-206
        _check(self._lib.incerto_builder_record_time_series(self._h, component.handle, identifier.handle,
                                                            sample_interval))
        return self

    def build(self, borrow_spawn_buffers: bool = False) -> "Simulation":          # :295-305
        """`borrow_spawn_buffers`: hand the spawners' staging arrays to the engine without a copy (they live until this
        call returns); the simulation then cannot be `reset`."""
        # spawners run in insertion order
        keep_all = []
        for sp in self._spawners:
            if isinstance(sp, DeviceSpawner):
                _check(self._lib.incerto_builder_add_device_spawner(
                    self._h, sp.kind, C.cast(C.pointer(sp.params), C.c_void_p), C.sizeof(sp.params)))
                continue
            spawner = Spawner()
            sp(spawner)
            spawner._flush_rows()
            for n, cols in spawner._calls:
                arr = (F.ColumnData * max(len(cols), 1))()
                keep = []
                for i, (comp, chunks) in enumerate(cols.items()):
                    arr[i].component = comp.handle
                    if chunks is None:
                        arr[i].data = None
                    else:
                        a = np.ascontiguousarray(chunks[0] if len(chunks) == 1 else np.concatenate(chunks, axis=0))
                        keep.append(a)
                        arr[i].data = a.ctypes.data_as(C.c_void_p)
                keep_all.append(keep)
                add = (self._lib.incerto_builder_add_entity_spawner_borrowed if borrow_spawn_buffers
                       else self._lib.incerto_builder_add_entity_spawner)
                _check(add(self._h, n, arr, len(cols)))
        out = C.c_void_p()
        h, self._h = self._h, None
        code = self._lib.incerto_builder_build(h, C.byref(out))   # consumes the builder, also on failure
        _check(code)
        return Simulation(out)


class TimeSeries:
    """src/types/times_series.rs:1-93 (values are copied out of the handle)."""

    def __init__(self, time: np.ndarray, values: list, sample_interval: int):
        self._time = time
        self._values = values
        self._interval = sample_interval

    def __len__(self): return len(self._values)
    def len(self): return len(self._values)
    def is_empty(self): return len(self._values) == 0
    def duration(self): return 0 if self.is_empty() else int(self._time[-1])
    def sample_interval(self): return self._interval
    def time(self): return [int(t) for t in self._time]
    def values(self): return list(self._values)
    def values_copied(self): return list(self._values)
    def enumerate(self): return list(zip(self.time(), self._values))
    def enumerate_copied(self): return self.enumerate()


class Simulation:
    """src/simulation.rs:17-259"""

    def __init__(self, handle: C.c_void_p):
        self._lib = F.load()
        self._h = handle

    def __del__(self):
        self.destroy()

    def destroy(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.incerto_sim_destroy(self._h)
            self._h = None

    def run(self, num_steps: int) -> None:                                      # :25-31
        _check(self._lib.incerto_sim_run(self._h, num_steps))

    def run_async(self, num_steps: int) -> None:
        _check(self._lib.incerto_sim_run_async(self._h, num_steps))

    def synchronize(self) -> None:
        _check(self._lib.incerto_sim_synchronize(self._h))

    def reset(self, replica: int) -> None:
        _check(self._lib.incerto_sim_reset(self._h, replica))

    def count(self, *filters) -> int:                                           # :163-173
        fs, keep = _filter_struct(filters)
        out = C.c_uint64()
        _check(self._lib.incerto_sim_count(self._h, C.byref(fs), C.byref(out)))
        return out.value

    def sample_aggregate(self, agg: Aggregate):                                 # :107-121
        d, keep = agg._desc()
        buf = (C.c_uint8 * 64)()
        _check(self._lib.incerto_sim_sample_aggregate(self._h, C.byref(d), buf, 64))
        return agg._decode(bytes(buf))

    def sample_aggregate_global(self, agg: Aggregate):
        """The aggregate over the entities of all ranks of the attached communicator (collective call)."""
        d, keep = agg._desc()
        buf = (C.c_uint8 * 64)()
        _check(self._lib.incerto_sim_sample_aggregate_global(self._h, C.byref(d), buf, 64))
        return agg._decode(bytes(buf))

    def sample_aggregate_filtered(self, agg: Aggregate, *filters):              # :136-152
        return self.sample_aggregate(agg.filtered(*filters))

    def sample(self, component: Component, identifier: Component, id_value):    # :43-65
        idv = np.asarray([id_value], dtype=identifier.np_dtype)
        out = np.zeros(component.lanes, dtype=component.np_dtype)
        _check(self._lib.incerto_sim_sample(self._h, component.handle, identifier.handle,
                                            idv.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out[0].item() if component.lanes == 1 else out

    def sample_single(self, component: Component):                              # :79-93
        out = np.zeros(component.lanes, dtype=component.np_dtype)
        _check(self._lib.incerto_sim_sample_single(self._h, component.handle, out.ctypes.data_as(C.c_void_p),
                                                   out.nbytes))
        return out[0].item() if component.lanes == 1 else out

    def get_aggregate_time_series(self, agg: Aggregate) -> TimeSeries:          # :226-258
        d, keep = agg._desc()
        tp = C.POINTER(C.c_uint64)()
        vp = C.c_void_p()
        n, rs, iv = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(self._lib.incerto_sim_get_aggregate_time_series(self._h, C.byref(d), C.byref(tp), C.byref(vp),
                                                               C.byref(n), C.byref(rs), C.byref(iv)))
        time = np.array([tp[i] for i in range(n.value)], dtype=np.uint64)
        raw = C.string_at(vp.value, n.value * rs.value) if n.value else b""
        values = [agg._decode(raw[i * rs.value:(i + 1) * rs.value]) for i in range(n.value)]
        return TimeSeries(time, values, iv.value)

    def get_aggregate_time_series_filtered(self, agg: Aggregate, *filters) -> TimeSeries:
        return self.get_aggregate_time_series(agg.filtered(*filters))

    def get_time_series(self, component: Component, identifier: Component, id_value) -> TimeSeries:   # :186-216
        idv = np.asarray([id_value], dtype=identifier.np_dtype)
        tp = C.POINTER(C.c_uint64)()
        vp = C.c_void_p()
        n, rs, iv = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(self._lib.incerto_sim_get_time_series(self._h, component.handle, identifier.handle,
                                                     idv.ctypes.data_as(C.c_void_p), C.byref(tp), C.byref(vp),
                                                     C.byref(n), C.byref(rs), C.byref(iv)))
        time = np.array([tp[i] for i in range(n.value)], dtype=np.uint64)
        raw = C.string_at(vp.value, n.value * rs.value) if n.value else b""
        vals = np.frombuffer(raw, dtype=component.np_dtype).reshape(n.value, component.lanes)
        values = [v[0].item() if component.lanes == 1 else v for v in vals]
        return TimeSeries(time, values, iv.value)

    # ---- SpatialGrid<T, C> views (src/plugins/spatial_grid.rs:244-389); entities reported by uid
    def grid_entities_at(self, grid: int, pos) -> list[int]:
        p = (C.c_int32 * 3)(*list(pos), *([0] * (3 - len(pos))))
        n = C.c_uint64()
        _check(self._lib.incerto_sim_grid_entities_at(self._h, grid, p, None, 0, C.byref(n)))
        out = (C.c_uint64 * max(n.value, 1))()
        _check(self._lib.incerto_sim_grid_entities_at(self._h, grid, p, out, n.value, C.byref(n)))
        return sorted(out[i] for i in range(n.value))

    def grid_neighbors_of(self, grid: int, pos, orthogonal: bool = False) -> list[int]:
        p = (C.c_int32 * 3)(*list(pos), *([0] * (3 - len(pos))))
        n = C.c_uint64()
        _check(self._lib.incerto_sim_grid_neighbors_of(self._h, grid, p, int(orthogonal), None, 0, C.byref(n)))
        out = (C.c_uint64 * max(n.value, 1))()
        _check(self._lib.incerto_sim_grid_neighbors_of(self._h, grid, p, int(orthogonal), out, n.value, C.byref(n)))
        return sorted(out[i] for i in range(n.value))

    def grid_is_empty(self, grid: int, pos) -> bool:
        p = (C.c_int32 * 3)(*list(pos), *([0] * (3 - len(pos))))
        out = C.c_uint32()
        _check(self._lib.incerto_sim_grid_is_empty(self._h, grid, p, C.byref(out)))
        return bool(out.value)

    def grid_num_entities(self, grid: int) -> int:
        out = C.c_uint64()
        _check(self._lib.incerto_sim_grid_num_entities(self._h, grid, C.byref(out)))
        return out.value

    def grid_position_of(self, grid: int, uid: int, dim: int = 2):
        out = (C.c_int32 * 3)()
        _check(self._lib.incerto_sim_grid_position_of(self._h, grid, uid, out))
        return tuple(out[i] for i in range(dim))

    def grid_in_bounds(self, grid: int, pos) -> bool:
        p = (C.c_int32 * 3)(*list(pos), *([0] * (3 - len(pos))))
        out = C.c_uint32()
        _check(self._lib.incerto_sim_grid_in_bounds(self._h, grid, p, C.byref(out)))
        return bool(out.value)

    # ---- introspection
    def step_number(self) -> int:
        out = C.c_uint64()
        _check(self._lib.incerto_sim_step_number(self._h, C.byref(out)))
        return out.value

    def num_entities(self, component: Component) -> int:
        out = C.c_uint64()
        _check(self._lib.incerto_sim_num_entities(self._h, component.handle, C.byref(out)))
        return out.value

    def read_component(self, component: Component, with_uids: bool = False, rows_of: Component | None = None):
        """Column contents of the live rows in row order (markers: 0/1 per live row of the tables holding `rows_of`)."""
        if component.dtype == F.MARKER:
            if rows_of is None:
                raise ValueError("read_component of a marker needs rows_of=<a data component of the same archetype>")
            n = self.num_entities(rows_of)
            out = np.zeros(n, dtype=np.uint8)
            _check(self._lib.incerto_sim_read_marker(self._h, component.handle, rows_of.handle,
                                                     out.ctypes.data_as(C.c_void_p), out.nbytes))
            if with_uids:
                _, uids = self.read_component(rows_of, with_uids=True)
                return out, uids
            return out
        else:
            n = self.num_entities(component)
            out = np.zeros((n, component.lanes), dtype=component.np_dtype)
        uids = np.zeros(n, dtype=np.uint64)
        _check(self._lib.incerto_sim_read_component(self._h, component.handle, out.ctypes.data_as(C.c_void_p),
                                                    out.nbytes, uids.ctypes.data_as(C.POINTER(C.c_uint64))))
        out = out[:, 0] if out.shape[1] == 1 else out
        return (out, uids) if with_uids else out

    def read_portfolio(self, portfolio: Component, rows_of: Component, num_stocks: int) -> np.ndarray:
        """Trader.portfolio (engine-managed): #shares per (live trader, stock); 0 = not owned."""
        n = self.num_entities(rows_of)
        out = np.zeros((n, num_stocks), dtype=np.uint8)
        _check(self._lib.incerto_sim_read_component(self._h, portfolio.handle, out.ctypes.data_as(C.c_void_p),
                                                    out.nbytes, None))
        return out

    def read_contact_history(self, contact_history: Component, rows_of: Component) -> list[set]:
        """ContactHistory.recent_contacts (engine-managed) of every live person as a set of (x, y)."""
        n = self.num_entities(rows_of)
        out = np.zeros((n, 21), dtype=np.uint32)
        _check(self._lib.incerto_sim_read_component(self._h, contact_history.handle, out.ctypes.data_as(C.c_void_p),
                                                    out.nbytes, None))
        return [set((int(k) & 0xffff, int(k) >> 16) for k in row[1:1 + row[0]]) for row in out]

    def last_run_stats(self) -> tuple[float, int]:
        ms, n = C.c_double(), C.c_uint64()
        _check(self._lib.incerto_sim_last_run_stats(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def set_profiling(self, enabled: bool) -> None:
        _check(self._lib.incerto_sim_set_profiling(self._h, int(enabled)))

    def kernel_stats(self) -> list[dict]:
        arr = (F.KernelStat * 64)()
        n = C.c_uint32()
        _check(self._lib.incerto_sim_kernel_stats(self._h, arr, 64, C.byref(n)))
        return [dict(name=arr[i].name.decode(), launches=arr[i].launches, device_ms=arr[i].device_ms,
                     algorithmic_bytes=arr[i].algorithmic_bytes) for i in range(min(n.value, 64))]

    def span_ms(self, last: "Simulation") -> float:
        """Device time from the start of this handle's last run to the end of `last`'s (handles driven concurrently
        with run_async on one GPU; both synchronised)."""
        out = C.c_double(0.0)
        _check(self._lib.incerto_sim_span_ms(self._h, last._h, C.byref(out)))
        return out.value

    def flush_l2(self) -> None:
        _check(self._lib.incerto_sim_flush_l2(self._h))

    def run_cold(self, num_steps: int) -> None:
        """`num_steps` steps, each from a flushed L2, enqueued back to back; `last_run_stats()` then holds the sum of
        the per-step device times (the flushes are not timed)."""
        _check(self._lib.incerto_sim_run_cold(self._h, num_steps))

    def halo_export(self, side: int) -> np.ndarray:
        """What the neighbour on `side` (0 lower, 1 upper) needs after the last step: the boundary column (+ overlay
        states) of a row-partitioned forest, or the mailbox (migrated aircraft + ghosts) of an x-partitioned airspace."""
        n = C.c_size_t()
        buf = np.zeros(1 << 16, dtype=np.uint8)
        code = self._lib.incerto_sim_halo_export(self._h, side, buf.ctypes.data_as(C.c_void_p), buf.nbytes, C.byref(n))
        if code == F.ERR_INVALID_ARGUMENT and n.value > buf.nbytes:      # the message is larger: the size came back
            buf = np.zeros(n.value, dtype=np.uint8)
            code = self._lib.incerto_sim_halo_export(self._h, side, buf.ctypes.data_as(C.c_void_p), buf.nbytes, C.byref(n))
        _check(code)
        return buf[:n.value].copy()

    def halo_import(self, side: int, data: np.ndarray) -> None:
        d = np.ascontiguousarray(data, dtype=np.uint8)
        _check(self._lib.incerto_sim_halo_import(self._h, side, d.ctypes.data_as(C.c_void_p), d.nbytes))

    def attach_comm(self, unique_id: bytes, rank: int, world: int) -> None:
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(self._lib.incerto_sim_attach_comm(self._h, buf, rank, world))

    def allreduce_u64(self, values: np.ndarray, op: int = 0) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.uint64)
        _check(self._lib.incerto_sim_allreduce_u64(self._h, v.ctypes.data_as(C.POINTER(C.c_uint64)), v.size, op))
        return v

    def allreduce_f64(self, values: np.ndarray, op: int = 0) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.float64)
        _check(self._lib.incerto_sim_allreduce_f64(self._h, v.ctypes.data_as(C.POINTER(C.c_double)), v.size, op))
        return v


def partition_range(width: int, rank: int, world: int) -> tuple[int, int]:
    b, e = C.c_int32(), C.c_int32()
    _check(F.load().incerto_partition_range(width, rank, world, C.byref(b), C.byref(e)))
    return b.value, e.value


def comm_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    _check(F.load().incerto_comm_unique_id(buf))
    return bytes(buf)
