"""ctypes binding of include/incerto_b200.h (the C ABI).  No compute happens here."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("INCERTO_B200_LIB") or os.path.join(_HERE, "lib", "libincerto_b200.so")

# ---- enums (mirror include/incerto_b200.h)
OK = 0
ERR_COMPONENT_DOES_NOT_EXIST = 1
ERR_SINGLE_NO_ENTITIES = 2
ERR_SINGLE_MULTIPLE_ENTITIES = 3
ERR_AGGREGATE_NO_ENTITIES = 4
ERR_TIME_SERIES_NOT_RECORDED = 5
ERR_ENTITY_IDENTIFIER_NOT_FOUND = 6
ERR_ENTITY_IDENTIFIER_NOT_UNIQUE = 7
ERR_TIME_SERIES_RECORDING_CONFLICT = 16
ERR_INVALID_ARGUMENT = 32
ERR_CUDA = 33
ERR_OUT_OF_BOUNDS = 34
ERR_NOT_COMPARABLE = 35
ERR_PERCENTILE_RANGE = 36
ERR_UNSUPPORTED = 37
ERR_NO_DEVICE = 38
ERR_COMM = 39

MARKER, U8, U16, U32, U64, I8, I16, I32, I64, F32, F64 = range(11)

SPAWN_COIN_TOSSERS, SPAWN_FOREST_GRID, SPAWN_PANDEMIC_PEOPLE, SPAWN_SPATIAL_POPULATION = 1, 2, 3, 4
SPAWN_TRADERS, SPAWN_STOCKS, SPAWN_AIRCRAFT = 5, 6, 7

SYS_ADD_CONST, SYS_ADD_COLUMN = 1, 2
SYS_COIN_TOSS = 10
SYS_FOREST_SPREAD, SYS_FOREST_BURN_PROGRESSION, SYS_FOREST_REGROWTH = 20, 21, 22
SYS_PANDEMIC_MOVE, SYS_PANDEMIC_INFECT, SYS_PANDEMIC_DIE, SYS_PANDEMIC_RECOVER = 30, 31, 32, 33
SYS_SPATIAL_INCUBATION, SYS_SPATIAL_TRANSMISSION, SYS_SPATIAL_RECOVERY_DEATH = 40, 41, 42
SYS_SPATIAL_CONTACT_HISTORY, SYS_SPATIAL_CONTACT_TRACING, SYS_SPATIAL_QUARANTINE, SYS_SPATIAL_MOVE = 43, 44, 45, 46
SYS_TRADERS_BUY, SYS_TRADERS_SELL, SYS_TRADERS_NET_WORTH, SYS_STOCKS_PRICE_CHANGE = 50, 51, 52, 53
SYS_AIR_CHECK_CONFLICTS, SYS_AIR_MOVE = 60, 61

AGG_SUM, AGG_COUNT, AGG_MINIMUM, AGG_MAXIMUM, AGG_MEAN, AGG_MEDIAN, AGG_PERCENTILE = 1, 2, 3, 4, 5, 6, 7
AGG_COIN_ODDS, AGG_FIRE_STATS, AGG_PANDEMIC_STATS, AGG_AIR_CONFLICTS = 16, 17, 18, 19

CELL_BURNED, CELL_HEALTHY, CELL_EMPTY = 0, 0x40, 0x80
DISEASE_HEALTHY, DISEASE_INFECTIOUS, DISEASE_RECOVERED = 0, 0x80, 0x81   # Exposed{n} = n

component_t = C.c_int32


class ColumnData(C.Structure):
    _fields_ = [("component", component_t), ("data", C.c_void_p)]


class SystemDesc(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("params", C.c_void_p), ("params_size", C.c_uint32)]


class Filter(C.Structure):
    _fields_ = [("with_", C.POINTER(component_t)), ("n_with", C.c_uint32),
                ("without", C.POINTER(component_t)), ("n_without", C.c_uint32)]


class AddConstParams(C.Structure):
    _fields_ = [("target", component_t), ("amount", C.c_int64), ("with_", component_t * 4), ("n_with", C.c_uint32)]


class AddColumnParams(C.Structure):
    _fields_ = [("target", component_t), ("source", component_t)]


class CoinTossParams(C.Structure):
    _fields_ = [("num_heads", component_t), ("num_tosses", component_t), ("odds_heads", C.c_double)]


class SpawnCoinTossersParams(C.Structure):
    _fields_ = [("num_heads", component_t), ("num_tosses", component_t), ("num_entities", C.c_uint64)]


class ForestParams(C.Structure):
    _fields_ = [("position", component_t), ("cell", component_t), ("probability", C.c_double),
                ("burn_duration", C.c_uint32)]


class SpawnForestParams(C.Structure):
    _fields_ = [("position", component_t), ("cell", component_t), ("width", C.c_int32), ("height", C.c_int32),
                ("initial_forest_density", C.c_double), ("burn_duration", C.c_uint32),
                ("initial_fire_count", C.c_uint32)]


class PandemicParams(C.Structure):
    _fields_ = [("person", component_t), ("position", component_t), ("infected", component_t),
                ("grid_size", C.c_uint32), ("probability", C.c_double)]


class SpawnPandemicParams(C.Structure):
    _fields_ = [("person", component_t), ("position", component_t), ("infected", component_t),
                ("initial_population", C.c_uint64), ("grid_size", C.c_uint32),
                ("chance_start_infected", C.c_double)]


class AirParams(C.Structure):
    _fields_ = [("position", component_t), ("target_altitude", component_t), ("airspace_size", C.c_int32),
                ("airspace_height", C.c_int32), ("chance_horizontal_move", C.c_double)]


class SpawnAircraftParams(C.Structure):
    _fields_ = [("position", component_t), ("target_altitude", component_t), ("num_aircraft", C.c_uint64),
                ("airspace_size", C.c_int32), ("airspace_height", C.c_int32)]


class SpatialParams(C.Structure):
    _fields_ = [("position", component_t), ("disease_state", component_t), ("social_distancing", component_t),
                ("quarantined", component_t), ("contact_history", component_t), ("grid_size", C.c_int32),
                ("infection_radius", C.c_int32), ("chance_infect_distance_1", C.c_double),
                ("chance_infect_distance_2", C.c_double), ("chance_recover", C.c_double), ("chance_die", C.c_double),
                ("incubation_period", C.c_uint32), ("crowding_threshold", C.c_uint32),
                ("contact_quarantine_duration", C.c_uint32), ("quarantine_center_x", C.c_int32),
                ("quarantine_center_y", C.c_int32), ("quarantine_radius", C.c_int32),
                ("contact_history_clear_above", C.c_uint32), ("chance_attempt_move", C.c_double)]


class SpawnSpatialParams(C.Structure):
    _fields_ = [("position", component_t), ("disease_state", component_t), ("social_distancing", component_t),
                ("quarantined", component_t), ("contact_history", component_t), ("initial_population", C.c_uint64),
                ("grid_size", C.c_int32), ("social_distance_compliance", C.c_double),
                ("chance_start_infected", C.c_double), ("incubation_period", C.c_uint32)]


class TradersParams(C.Structure):
    _fields_ = [("trader_id", component_t), ("cash", component_t), ("portfolio", component_t),
                ("net_worth", component_t), ("stock_id", component_t), ("stock_price", component_t),
                ("stock_delisted", component_t), ("chance_buy", C.c_double), ("chance_sell", C.c_double),
                ("buy_shares_min", C.c_uint32), ("buy_shares_max", C.c_uint32), ("price_change_mean", C.c_double),
                ("price_change_std_dev", C.c_double)]


class SpawnTradersParams(C.Structure):
    _fields_ = [("trader_id", component_t), ("cash", component_t), ("portfolio", component_t),
                ("net_worth", component_t), ("num_traders", C.c_uint64), ("initial_cash", C.c_double),
                ("num_stocks", C.c_uint32)]


class SpawnStocksParams(C.Structure):
    _fields_ = [("stock_id", component_t), ("stock_price", component_t), ("stock_delisted", component_t),
                ("num_stocks", C.c_uint32), ("initial_price", C.c_double)]


class FireStats(C.Structure):
    _fields_ = [("healthy_count", C.c_uint64), ("burning_count", C.c_uint64), ("burned_count", C.c_uint64),
                ("empty_count", C.c_uint64), ("total_cells", C.c_uint64)]


class PandemicStats(C.Structure):
    _fields_ = [("healthy_count", C.c_uint64), ("exposed_count", C.c_uint64), ("infectious_count", C.c_uint64),
                ("recovered_count", C.c_uint64), ("dead_count", C.c_uint64), ("total_population", C.c_uint64)]


class AggregateDesc(C.Structure):
    _fields_ = [("component", component_t), ("component2", component_t), ("kind", C.c_uint32),
                ("percentile", C.c_uint32), ("param", C.c_uint64), ("filter", Filter)]


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char_p), ("launches", C.c_uint64), ("device_ms", C.c_double),
                ("algorithmic_bytes", C.c_double)]


# every symbol include/incerto_b200.h declares (tests/test_abi.py checks this list against the header)
SYMBOLS = [
    "incerto_last_error_message", "incerto_abi_version", "incerto_builder_register_component",
    "incerto_builder_new", "incerto_builder_free", "incerto_builder_set_seed", "incerto_builder_set_device",
    "incerto_builder_set_row_partition", "incerto_builder_add_entity_spawner", "incerto_builder_add_entity_spawner_borrowed",
    "incerto_builder_add_device_spawner",
    "incerto_builder_add_systems", "incerto_builder_add_spatial_grid",
    "incerto_builder_record_aggregate_time_series", "incerto_builder_record_time_series", "incerto_builder_build",
    "incerto_sim_run", "incerto_sim_run_async", "incerto_sim_synchronize", "incerto_sim_destroy", "incerto_sim_reset",
    "incerto_sim_count", "incerto_sim_sample_aggregate", "incerto_sim_sample_aggregate_global", "incerto_sim_sample", "incerto_sim_sample_single",
    "incerto_sim_get_aggregate_time_series", "incerto_sim_get_time_series", "incerto_sim_grid_entities_at",
    "incerto_sim_grid_neighbors_of", "incerto_sim_grid_is_empty", "incerto_sim_grid_num_entities",
    "incerto_sim_grid_position_of", "incerto_sim_grid_in_bounds", "incerto_sim_step_number",
    "incerto_sim_num_entities", "incerto_sim_read_component", "incerto_sim_read_marker", "incerto_sim_last_run_stats",
    "incerto_sim_span_ms",
    "incerto_sim_set_profiling", "incerto_sim_kernel_stats", "incerto_sim_flush_l2", "incerto_sim_run_cold", "incerto_comm_unique_id",
    "incerto_sim_attach_comm", "incerto_sim_allreduce_u64", "incerto_sim_allreduce_f64", "incerto_sim_halo_export",
    "incerto_sim_halo_import", "incerto_partition_range",
]

_lib = None


def load() -> C.CDLL:
    """Load the CUDA engine.  Fails loudly when it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build the CUDA engine first (python -m incerto_b200._build); "
            "incerto_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.incerto_last_error_message.restype = C.c_char_p
    lib.incerto_abi_version.restype = C.c_uint32
    lib.incerto_builder_free.restype = None
    lib.incerto_sim_destroy.restype = None
    vp = C.c_void_p
    lib.incerto_builder_new.argtypes = [C.POINTER(vp)]
    lib.incerto_builder_free.argtypes = [vp]
    lib.incerto_builder_set_seed.argtypes = [vp, C.c_uint64, C.c_uint32]
    lib.incerto_builder_set_device.argtypes = [vp, C.c_int32]
    lib.incerto_builder_set_row_partition.argtypes = [vp, C.c_uint32, C.c_uint32]
    lib.incerto_builder_register_component.argtypes = [vp, C.c_char_p, C.c_int, C.c_uint32, C.POINTER(component_t)]
    lib.incerto_builder_add_entity_spawner.argtypes = [vp, C.c_uint64, C.POINTER(ColumnData), C.c_uint32]
    lib.incerto_builder_add_entity_spawner_borrowed.argtypes = [vp, C.c_uint64, C.POINTER(ColumnData), C.c_uint32]
    lib.incerto_builder_add_device_spawner.argtypes = [vp, C.c_uint32, vp, C.c_uint32]
    lib.incerto_builder_add_systems.argtypes = [vp, C.POINTER(SystemDesc), C.c_uint32]
    lib.incerto_builder_add_spatial_grid.argtypes = [vp, C.c_uint32, C.POINTER(C.c_int32), component_t, component_t,
                                                     C.POINTER(C.c_int32)]
    lib.incerto_builder_record_aggregate_time_series.argtypes = [vp, C.POINTER(AggregateDesc), C.c_uint64]
    lib.incerto_builder_record_time_series.argtypes = [vp, component_t, component_t, C.c_uint64]
    lib.incerto_builder_build.argtypes = [vp, C.POINTER(vp)]
    lib.incerto_sim_run.argtypes = [vp, C.c_uint64]
    lib.incerto_sim_run_async.argtypes = [vp, C.c_uint64]
    lib.incerto_sim_synchronize.argtypes = [vp]
    lib.incerto_sim_destroy.argtypes = [vp]
    lib.incerto_sim_reset.argtypes = [vp, C.c_uint32]
    lib.incerto_sim_count.argtypes = [vp, C.POINTER(Filter), C.POINTER(C.c_uint64)]
    lib.incerto_sim_sample_aggregate.argtypes = [vp, C.POINTER(AggregateDesc), vp, C.c_size_t]
    lib.incerto_sim_sample_aggregate_global.argtypes = [vp, C.POINTER(AggregateDesc), vp, C.c_size_t]
    lib.incerto_sim_sample.argtypes = [vp, component_t, component_t, vp, vp, C.c_size_t]
    lib.incerto_sim_sample_single.argtypes = [vp, component_t, vp, C.c_size_t]
    lib.incerto_sim_get_aggregate_time_series.argtypes = [vp, C.POINTER(AggregateDesc), C.POINTER(C.POINTER(C.c_uint64)),
                                                          C.POINTER(vp), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                                          C.POINTER(C.c_uint64)]
    lib.incerto_sim_get_time_series.argtypes = [vp, component_t, component_t, vp, C.POINTER(C.POINTER(C.c_uint64)),
                                                C.POINTER(vp), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                                C.POINTER(C.c_uint64)]
    lib.incerto_sim_grid_entities_at.argtypes = [vp, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.c_uint64,
                                                 C.POINTER(C.c_uint64)]
    lib.incerto_sim_grid_neighbors_of.argtypes = [vp, C.c_int32, C.POINTER(C.c_int32), C.c_uint32,
                                                  C.POINTER(C.c_uint64), C.c_uint64, C.POINTER(C.c_uint64)]
    lib.incerto_sim_grid_is_empty.argtypes = [vp, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_uint32)]
    lib.incerto_sim_grid_num_entities.argtypes = [vp, C.c_int32, C.POINTER(C.c_uint64)]
    lib.incerto_sim_grid_position_of.argtypes = [vp, C.c_int32, C.c_uint64, C.POINTER(C.c_int32)]
    lib.incerto_sim_grid_in_bounds.argtypes = [vp, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_uint32)]
    lib.incerto_sim_step_number.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.incerto_sim_num_entities.argtypes = [vp, component_t, C.POINTER(C.c_uint64)]
    lib.incerto_sim_read_component.argtypes = [vp, component_t, vp, C.c_size_t, C.POINTER(C.c_uint64)]
    lib.incerto_sim_read_marker.argtypes = [vp, component_t, component_t, vp, C.c_size_t]
    lib.incerto_sim_last_run_stats.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    lib.incerto_sim_set_profiling.argtypes = [vp, C.c_uint32]
    lib.incerto_sim_kernel_stats.argtypes = [vp, C.POINTER(KernelStat), C.c_uint32, C.POINTER(C.c_uint32)]
    lib.incerto_sim_flush_l2.argtypes = [vp]
    lib.incerto_sim_run_cold.argtypes = [vp, C.c_uint64]
    lib.incerto_sim_span_ms.argtypes = [vp, vp, C.POINTER(C.c_double)]
    lib.incerto_comm_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    lib.incerto_sim_attach_comm.argtypes = [vp, C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32]
    lib.incerto_sim_allreduce_u64.argtypes = [vp, C.POINTER(C.c_uint64), C.c_uint32, C.c_uint32]
    lib.incerto_sim_allreduce_f64.argtypes = [vp, C.POINTER(C.c_double), C.c_uint32, C.c_uint32]
    lib.incerto_sim_halo_export.argtypes = [vp, C.c_uint32, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.incerto_sim_halo_import.argtypes = [vp, C.c_uint32, vp, C.c_size_t]
    lib.incerto_partition_range.argtypes = [C.c_int32, C.c_uint32, C.c_uint32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    _lib = lib
    return lib


def last_error() -> str:
    return load().incerto_last_error_message().decode("utf-8", "replace")
