"""ORACLE — test infrastructure, not product code.

Pure-Python restatement of the incerto library (src/) on a miniature Bevy: enough of the ECS
semantics the reference relies on to run arbitrary user systems the way
`Simulation::run` does.  Used (a) to pin the oracle against the reference's own RNG-free
tests (tests/test_oracle_kats.py ports every case of /tests/*.rs) and (b) as a second,
independent statement of the example systems at tiny sizes (oracle/examples_ref.py).

Reference files followed:
  src/simulation_builder.rs:25-306   SimulationBuilder
  src/simulation.rs:17-259           Simulation
  src/spawner.rs:3-12                Spawner
  src/plugins/step_number.rs:6-23    StepNumber (+1 in Last)
  src/plugins/time_series.rs:42-180  aggregate and per-entity time series (PostUpdate)
  src/plugins/spatial_grid.rs:1-482  SpatialGrid (PreUpdate maintenance)
  src/util/aggregate.rs:73-192       Minimum / Maximum / Mean / Median / Percentile
  src/types/times_series.rs:1-93     TimeSeries view
  src/error.rs:1-73                  errors

Bevy semantics restated [unverified against bevy source: not vendored]:
  * schedules per update: PreUpdate -> Update -> PostUpdate -> Last;
  * Update systems run in insertion order (one legal linearisation);
  * Commands are deferred to the end of the schedule that issued them
    (chained systems get a sync point in between, time_series.rs:175-178);
  * direct component writes are immediate;
  * query iteration order = spawn order; components are plain Python objects.
"""
from __future__ import annotations

import math
from typing import Any, Callable


# ------------------------------------------------------------------ errors
class SamplingError(Exception):
    def __init__(self, kind: str):
        self.kind = kind
        super().__init__(kind)


class BuilderError(Exception):
    def __init__(self, kind: str):
        self.kind = kind
        super().__init__(kind)


# ------------------------------------------------------------------- world
class World:
    def __init__(self):
        self.entities: dict[int, dict[type, Any]] = {}
        self.next_id = 0
        self.registered: set[type] = set()    # component ids known to the world (try_query -> None otherwise)
        self.resources: dict[Any, Any] = {}
        self.removed: dict[type, list[int]] = {}   # RemovedComponents<T>

    def spawn(self, *components) -> int:
        e = self.next_id
        self.next_id += 1
        self.entities[e] = {}
        for c in components:
            self.entities[e][type(c)] = c
            self.registered.add(type(c))
        return e

    def query(self, *types, with_=(), without=()):
        """[(entity, comp...)] in spawn order; marks nothing as changed."""
        for t in list(types) + list(with_) + list(without):
            self.registered.add(t)
        out = []
        for e, comps in self.entities.items():
            if all(t in comps for t in types) and all(t in comps for t in with_) and not any(t in comps for t in without):
                out.append((e, *[comps[t] for t in types]))
        return out

    def try_query(self, *types, with_=(), without=()):
        for t in list(types) + list(with_) + list(without):
            if t not in self.registered:
                return None
        return self.query(*types, with_=with_, without=without)

    def get(self, e: int, t: type):
        comps = self.entities.get(e)
        return None if comps is None else comps.get(t)


class Commands:
    """Deferred structural edits (insert / remove / despawn / try_insert)."""

    def __init__(self):
        self.queue: list[tuple] = []

    def insert(self, e: int, component) -> None:
        self.queue.append(("insert", e, component))

    def try_insert(self, e: int, component) -> None:
        self.queue.append(("try_insert", e, component))

    def remove(self, e: int, t: type) -> None:
        self.queue.append(("remove", e, t))

    def despawn(self, e: int) -> None:
        self.queue.append(("despawn", e))

    def apply(self, world: World) -> None:
        for op in self.queue:
            if op[0] in ("insert", "try_insert"):
                _, e, c = op
                if e not in world.entities:
                    if op[0] == "insert":
                        raise RuntimeError("insert on a despawned entity panics in bevy")
                    continue
                world.entities[e][type(c)] = c
                world.registered.add(type(c))
            elif op[0] == "remove":
                _, e, t = op
                if e in world.entities and t in world.entities[e]:
                    del world.entities[e][t]
                    world.removed.setdefault(t, []).append(e)
            else:
                _, e = op
                if e in world.entities:
                    for t in world.entities[e]:
                        world.removed.setdefault(t, []).append(e)
                    del world.entities[e]
        self.queue.clear()


class Ctx:
    """What a system sees: the world, a command queue, the resources."""

    def __init__(self, world: World, commands: Commands):
        self.world = world
        self.commands = commands

    def res(self, key):
        return self.world.resources[key]


PRE_UPDATE, UPDATE, POST_UPDATE, LAST = "PreUpdate", "Update", "PostUpdate", "Last"


class App:
    def __init__(self):
        self.world = World()
        # each schedule: list of chains; a chain is a list of systems with a sync point between them
        self.schedules: dict[str, list[list[Callable[[Ctx], None]]]] = {PRE_UPDATE: [], UPDATE: [], POST_UPDATE: [], LAST: []}

    def add_systems(self, schedule: str, *systems, chain: bool = False):
        if chain:
            self.schedules[schedule].append(list(systems))
        else:
            for s in systems:
                self.schedules[schedule].append([s])

    def update(self):
        for name in (PRE_UPDATE, UPDATE, POST_UPDATE, LAST):
            cmds = Commands()
            for chain in self.schedules[name]:
                for i, system in enumerate(chain):
                    if i > 0:
                        cmds.apply(self.world)    # sync point inserted by .chain()
                    system(Ctx(self.world, cmds))
            cmds.apply(self.world)                # end of schedule


# ---------------------------------------------------------------- plugins
class StepNumber:
    def __init__(self):
        self.value = 0


def step_number_plugin(app: App):
    """src/plugins/step_number.rs:10-23"""
    app.world.resources[StepNumber] = StepNumber()

    def step_counter_increment(ctx: Ctx):
        ctx.res(StepNumber).value += 1

    app.add_systems(LAST, step_counter_increment)


class GridBounds:
    def __init__(self, min_, max_):
        self.min, self.max = tuple(min_), tuple(max_)

    def contains(self, pos) -> bool:                         # spatial_grid.rs:208-242
        assert all(a <= b for a, b in zip(self.min, self.max))
        return all(lo <= p <= hi for p, lo, hi in zip(pos, self.min, self.max))


class GridPosition:
    """spatial_grid.rs:117-190 (2 or 3 integer coordinates)"""

    def __init__(self, *coords):
        self.coords = tuple(int(c) for c in coords)

    def x(self): return self.coords[0]
    def y(self): return self.coords[1]
    def z(self): return self.coords[2]
    def __eq__(self, o): return isinstance(o, GridPosition) and self.coords == o.coords
    def __hash__(self): return hash(self.coords)
    def __repr__(self): return f"GridPosition{self.coords}"

    def neighbors(self):                                     # :55-64, :76-97
        if len(self.coords) == 2:
            dirs = [(-1, -1), (0, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (0, 1), (1, 1)]
        else:
            dirs = [(dx, dy, dz) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1) if (dx, dy, dz) != (0, 0, 0)]
        return [GridPosition(*[c + d for c, d in zip(self.coords, dd)]) for dd in dirs]

    def neighbors_orthogonal(self):                          # :65-69, :99-103
        if len(self.coords) == 2:
            dirs = [(0, -1), (-1, 0), (1, 0), (0, 1)]
        else:
            dirs = [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]
        return [GridPosition(*[c + d for c, d in zip(self.coords, dd)]) for dd in dirs]


class SpatialGrid:
    """spatial_grid.rs:194-389"""

    def __init__(self, bounds: GridBounds | None):
        self.position_to_entities: dict[GridPosition, set[int]] = {}
        self.entity_to_position: dict[int, GridPosition] = {}
        self.bounds = bounds

    def in_bounds(self, position: GridPosition) -> bool:
        return self.bounds is None or self.bounds.contains(position.coords)

    def insert_or_update(self, entity: int, position: GridPosition):
        assert self.in_bounds(position), f"entity at position {position} outside spatial grid bounds"
        self.remove(entity)
        self.position_to_entities.setdefault(position, set()).add(entity)
        self.entity_to_position[entity] = position

    def remove(self, entity: int):
        position = self.entity_to_position.pop(entity, None)
        if position is None:
            return None
        s = self.position_to_entities[position]
        s.discard(entity)
        if not s:
            del self.position_to_entities[position]
        return position

    def entities_at(self, position: GridPosition):
        return list(self.position_to_entities.get(position, ()))

    def position_of(self, entity: int):
        return self.entity_to_position.get(entity)

    def neighbors_of(self, position: GridPosition):
        out = []
        for q in position.neighbors():
            if self.bounds is None or self.bounds.contains(q.coords):
                out.extend(self.entities_at(q))
        return out

    def orthogonal_neighbors_of(self, position: GridPosition):
        out = []
        for q in position.neighbors_orthogonal():
            if self.bounds is None or self.bounds.contains(q.coords):
                out.extend(self.entities_at(q))
        return out

    def is_empty(self, position: GridPosition) -> bool:
        return not self.position_to_entities.get(position)

    def num_entities(self) -> int:
        return len(self.entity_to_position)


def spatial_grid_plugin(app: App, component: type, bounds: GridBounds | None, key=None):
    """spatial_grid.rs:410-455: resource + (update, cleanup).chain() in PreUpdate."""
    key = key or (SpatialGrid, component)
    grid = SpatialGrid(bounds)
    app.world.resources[key] = grid
    seen: dict[int, GridPosition] = {}   # stands in for Changed<GridPosition<T>>

    def spatial_grid_update_system(ctx: Ctx):
        for e, pos in ctx.world.query(GridPosition, with_=(component,)):
            if seen.get(e) != pos:
                grid.insert_or_update(e, GridPosition(*pos.coords))
                seen[e] = GridPosition(*pos.coords)

    def spatial_grid_cleanup_system(ctx: Ctx):
        for e in ctx.world.removed.pop(GridPosition, []):
            grid.remove(e)
            seen.pop(e, None)

    app.add_systems(PRE_UPDATE, spatial_grid_update_system, spatial_grid_cleanup_system, chain=True)


class TimeSeriesData:
    def __init__(self, sample_interval: int):
        self.values: list = []
        self.time: list[int] = []
        self.sample_interval = sample_interval


class TimeSeries:
    """src/types/times_series.rs"""

    def __init__(self, data: TimeSeriesData):
        self._values = list(data.values)
        self._time = list(data.time)
        self._interval = data.sample_interval

    def len(self): return len(self._values)
    def __len__(self): return len(self._values)
    def is_empty(self): return not self._values
    def duration(self): return 0 if self.is_empty() else self._time[-1]
    def sample_interval(self): return self._interval
    def time(self): return list(self._time)
    def values(self): return list(self._values)
    def values_copied(self): return list(self._values)
    def enumerate(self): return list(zip(self._time, self._values))


# ------------------------------------------------------------- aggregators
def _ordered(values):
    for v in values:
        if isinstance(v, float) and math.isnan(v):
            raise ValueError("error during aggregate sampling, cannot compare values")   # aggregate.rs:227-231
    return values


def minimum(values): return min(_ordered(values))                     # aggregate.rs:73-88
def maximum(values): return max(_ordered(values))                     # :90-105


def percentile(values, p: int):                                       # :107-135
    assert 0 <= p <= 100
    s = sorted(_ordered(values))
    idx = math.floor((p / 100.0) * float(len(s)))
    return s[int(idx)]   # IndexError for p == 100, like the reference's out-of-range panic


def median(values):                                                   # :137-155
    s = sorted(_ordered(values))
    return s[len(s) // 2]


def mean(values, integer: bool):                                      # :157-192
    total = sum(values)
    return int(total) // len(values) if integer else total / float(len(values))


# ---------------------------------------------------------------- builder
class Spawner:
    def __init__(self, world: World):
        self.world = world

    def spawn(self, *components):
        self.world.spawn(*components)


class SimulationBuilder:
    def __init__(self):
        self.app = App()
        step_number_plugin(self.app)
        self.app.update()                       # simulation_builder.rs:50: StepNumber becomes 1
        self.spawners: list[Callable[[Spawner], None]] = []

    def add_systems(self, *systems, chain: bool = False):
        self.app.add_systems(UPDATE, *systems, chain=chain)
        return self

    def add_spatial_grid(self, component: type, bounds: GridBounds | None):
        spatial_grid_plugin(self.app, component, bounds)
        return self

    def add_entity_spawner(self, fn):
        self.spawners.append(fn)
        return self

    def add_resource(self, key, value):
        self.app.world.resources[key] = value
        return self

    def record_aggregate_time_series(self, component: type, out_key, sample_aggregate, sample_interval: int,
                                     with_=(), without=()):
        """(C, F, O) keyed resource; `sample_aggregate(list_of_components) -> O` (time_series.rs:42-102)."""
        assert sample_interval > 0
        key = (TimeSeriesData, component, (tuple(with_), tuple(without)), out_key)
        if key in self.app.world.resources:
            raise BuilderError("TimeSeriesRecordingConflict")
        data = TimeSeriesData(sample_interval)
        self.app.world.resources[key] = data

        def time_series_sample(ctx: Ctx):
            step = ctx.res(StepNumber).value
            if step % data.sample_interval == 0:
                comps = [c for _, c in ctx.world.query(component, with_=with_, without=without)]
                if comps:
                    data.values.append(sample_aggregate(comps))
                    data.time.append(step)

        self.app.add_systems(POST_UPDATE, time_series_sample)
        return self

    def record_time_series(self, component: type, identifier: type, out_key, sample, sample_interval: int):
        """time_series.rs:104-180"""
        assert sample_interval > 0
        key = ("SampleInterval", component, identifier, out_key)
        if key in self.app.world.resources:
            raise BuilderError("TimeSeriesRecordingConflict")
        self.app.world.resources[key] = sample_interval
        tag = type(f"TimeSeriesData_{component.__name__}_{identifier.__name__}_{out_key}", (TimeSeriesData,), {})
        self.app.world.resources[("series_type", component, identifier, out_key)] = tag
        added: set[int] = set()

        def time_series_init(ctx: Ctx):
            for e, _ in ctx.world.query(component):
                if e not in added:
                    added.add(e)
                    ctx.commands.insert(e, tag(sample_interval))

        def time_series_sample(ctx: Ctx):
            step = ctx.res(StepNumber).value
            for _, c, series in ctx.world.query(component, tag):
                if step % series.sample_interval == 0:
                    series.values.append(sample(c))
                    series.time.append(step)

        self.app.add_systems(POST_UPDATE, time_series_init, time_series_sample, chain=True)
        return self

    def build(self) -> "Simulation":
        spawner = Spawner(self.app.world)
        for fn in self.spawners:
            fn(spawner)
        return Simulation(self.app)


class Simulation:
    def __init__(self, app: App):
        self.app = app

    def run(self, num_steps: int):
        for _ in range(num_steps):
            self.app.update()

    def sample(self, component: type, identifier: type, id_value, sample=lambda c: c):
        q = self.app.world.try_query(component, identifier)
        if q is None:
            raise SamplingError("ComponentDoesNotExist")
        hits = [c for _, c, i in q if i == id_value]
        if not hits:
            raise SamplingError("EntityIdentifierNotFound")
        if len(hits) > 1:
            raise SamplingError("EntityIdentifierNotUnique")
        return sample(hits[0])

    def sample_single(self, component: type, sample=lambda c: c):
        q = self.app.world.try_query(component)
        if q is None:
            raise SamplingError("ComponentDoesNotExist")
        if not q:
            raise SamplingError("SingleNoEntities")
        if len(q) > 1:
            raise SamplingError("SingleMultipleEntities")
        return sample(q[0][1])

    def sample_aggregate(self, component: type, sample_aggregate, with_=(), without=()):
        q = self.app.world.try_query(component, with_=with_, without=without)
        if q is None:
            raise SamplingError("ComponentDoesNotExist")
        comps = [c for _, c in q]
        if not comps:
            raise SamplingError("AggregateNoEntities")
        return sample_aggregate(comps)

    def count(self, with_=(), without=()) -> int:
        q = self.app.world.try_query(with_=with_, without=without)
        if q is None:
            raise SamplingError("ComponentDoesNotExist")
        return len(q)

    def get_aggregate_time_series(self, component: type, out_key, with_=(), without=()) -> TimeSeries:
        key = (TimeSeriesData, component, (tuple(with_), tuple(without)), out_key)
        data = self.app.world.resources.get(key)
        if data is None:
            raise SamplingError("TimeSeriesNotRecorded")
        return TimeSeries(data)

    def get_time_series(self, component: type, identifier: type, out_key, id_value) -> TimeSeries:
        tag = self.app.world.resources.get(("series_type", component, identifier, out_key))
        q = None if tag is None else self.app.world.try_query(tag, identifier)
        if q is None:
            raise SamplingError("ComponentDoesNotExist")
        hits = [s for _, s, i in q if i == id_value]
        if not hits:
            raise SamplingError("EntityIdentifierNotFound")
        if len(hits) > 1:
            raise SamplingError("EntityIdentifierNotUnique")
        return TimeSeries(hits[0])
