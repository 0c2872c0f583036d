"""ORACLE — test infrastructure, not product code.  ctypes loader of oracle/_build/liboracle.so."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc"))]
    newest = max(os.path.getmtime(s) for s in srcs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return LIB


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(LIB)
    vp = C.c_void_p
    u64p, u32p, i32p, u8p = (C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_uint8))
    lib.oracle_philox4x32_10.argtypes = [u32p, u32p, u32p]
    lib.oracle_philox4x32_10.restype = None
    lib.oracle_threshold.argtypes = [C.c_double]
    lib.oracle_threshold.restype = C.c_uint64
    lib.oracle_below.argtypes = [C.c_uint32, C.c_uint32]
    lib.oracle_below.restype = C.c_uint32
    lib.oracle_lottery256.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, u8p]
    lib.oracle_lottery256.restype = C.c_uint32
    lib.oracle_coin_toss_run.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_double, C.c_uint64, C.c_uint64, u64p, u64p]
    lib.oracle_coin_toss_run.restype = C.c_double
    lib.oracle_forest_new.argtypes = [C.c_uint64, C.c_uint32, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_uint32,
                                      C.c_double, C.c_uint32, C.c_uint32]
    lib.oracle_forest_new.restype = vp
    lib.oracle_forest_free.argtypes = [vp]
    lib.oracle_forest_free.restype = None
    lib.oracle_forest_spawn.argtypes = [vp]
    lib.oracle_forest_spawn.restype = None
    lib.oracle_forest_spawn_explicit.argtypes = [vp, C.c_uint64, i32p, u8p]
    lib.oracle_forest_spawn_explicit.restype = None
    lib.oracle_forest_run.argtypes = [vp, C.c_uint64]
    lib.oracle_forest_run.restype = None
    lib.oracle_forest_num_entities.argtypes = [vp]
    lib.oracle_forest_num_entities.restype = C.c_uint64
    lib.oracle_forest_cells.argtypes = [vp, u8p, i32p]
    lib.oracle_forest_cells.restype = None
    lib.oracle_forest_stats.argtypes = [vp, u64p]
    lib.oracle_forest_stats.restype = None
    lib.oracle_forest_series_len.argtypes = [vp]
    lib.oracle_forest_series_len.restype = C.c_uint64
    lib.oracle_forest_series.argtypes = [vp, u64p]
    lib.oracle_forest_series.restype = None
    lib.oracle_forest_bench.argtypes = [C.c_uint64, C.c_int32, C.c_int32, C.c_uint64, C.c_uint32]
    lib.oracle_forest_bench.restype = C.c_uint64
    u16p = C.POINTER(C.c_uint16)
    lib.oracle_pandemic_new.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_double, C.c_double, C.c_double,
                                        C.c_double, C.c_uint32, C.c_uint32]
    lib.oracle_pandemic_new.restype = vp
    lib.oracle_pandemic_free.argtypes = [vp]
    lib.oracle_pandemic_free.restype = None
    lib.oracle_pandemic_spawn.argtypes = [vp]
    lib.oracle_pandemic_spawn.restype = None
    lib.oracle_pandemic_spawn_explicit.argtypes = [vp, C.c_uint64, u16p, u8p]
    lib.oracle_pandemic_spawn_explicit.restype = None
    lib.oracle_pandemic_run.argtypes = [vp, C.c_uint64]
    lib.oracle_pandemic_run.restype = None
    lib.oracle_pandemic_num_slots.argtypes = [vp]
    lib.oracle_pandemic_num_slots.restype = C.c_uint64
    lib.oracle_pandemic_state.argtypes = [vp, u16p, u8p, u8p]
    lib.oracle_pandemic_state.restype = None
    lib.oracle_pandemic_counts.argtypes = [vp, u64p]
    lib.oracle_pandemic_counts.restype = None
    lib.oracle_air_new.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_int32, C.c_int32, C.c_double, C.c_uint32]
    lib.oracle_air_new.restype = vp
    lib.oracle_air_free.argtypes = [vp]
    lib.oracle_air_free.restype = None
    lib.oracle_air_spawn.argtypes = [vp]
    lib.oracle_air_spawn.restype = None
    lib.oracle_air_spawn_explicit.argtypes = [vp, C.c_uint64, i32p, i32p]
    lib.oracle_air_spawn_explicit.restype = None
    lib.oracle_air_run.argtypes = [vp, C.c_uint64]
    lib.oracle_air_run.restype = C.c_int
    lib.oracle_air_num.argtypes = [vp]
    lib.oracle_air_num.restype = C.c_uint64
    lib.oracle_air_state.argtypes = [vp, i32p, i32p]
    lib.oracle_air_state.restype = None
    lib.oracle_air_series.argtypes = [vp, u64p, C.c_uint64]
    lib.oracle_air_series.restype = C.c_uint64
    f64p = C.POINTER(C.c_double)
    lib.oracle_spatial_new.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_int32, f64p, i32p, C.c_uint32]
    lib.oracle_spatial_new.restype = vp
    lib.oracle_spatial_free.argtypes = [vp]
    lib.oracle_spatial_free.restype = None
    lib.oracle_spatial_spawn.argtypes = [vp]
    lib.oracle_spatial_spawn.restype = None
    lib.oracle_spatial_spawn_explicit.argtypes = [vp, C.c_uint64, i32p, u8p, u8p]
    lib.oracle_spatial_spawn_explicit.restype = None
    lib.oracle_spatial_run.argtypes = [vp, C.c_uint64]
    lib.oracle_spatial_run.restype = C.c_int
    lib.oracle_spatial_num_slots.argtypes = [vp]
    lib.oracle_spatial_num_slots.restype = C.c_uint64
    lib.oracle_spatial_state.argtypes = [vp, i32p, u8p, u8p, u8p, u32p, u32p]
    lib.oracle_spatial_state.restype = None
    lib.oracle_spatial_stats.argtypes = [vp, u64p]
    lib.oracle_spatial_stats.restype = None
    lib.oracle_spatial_series.argtypes = [vp, u64p, C.c_uint64]
    lib.oracle_spatial_series.restype = C.c_uint64
    lib.oracle_traders_new.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_double, C.c_double, C.c_double,
                                       C.c_uint32, C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_uint32]
    lib.oracle_traders_new.restype = vp
    lib.oracle_traders_free.argtypes = [vp]
    lib.oracle_traders_free.restype = None
    lib.oracle_traders_set_prices.argtypes = [vp, f64p]
    lib.oracle_traders_set_prices.restype = None
    lib.oracle_traders_run.argtypes = [vp, C.c_uint64]
    lib.oracle_traders_run.restype = None
    lib.oracle_traders_state.argtypes = [vp, f64p, f64p, u8p, f64p, u8p]
    lib.oracle_traders_state.restype = None
    lib.oracle_traders_series.argtypes = [vp, C.c_uint64, f64p, C.c_uint64]
    lib.oracle_traders_series.restype = C.c_uint64
    lib.oracle_det_log.argtypes = [C.c_double]
    lib.oracle_det_log.restype = C.c_double
    lib.oracle_det_cos2pi.argtypes = [C.c_double]
    lib.oracle_det_cos2pi.restype = C.c_double
    lib.oracle_det_normal.argtypes = [C.c_uint32] * 4
    lib.oracle_det_normal.restype = C.c_double
    for name, ct in (("u8", C.c_uint8), ("u16", C.c_uint16), ("u32", C.c_uint32), ("u64", C.c_uint64),
                     ("i8", C.c_int8), ("i16", C.c_int16), ("i32", C.c_int32), ("i64", C.c_int64),
                     ("f32", C.c_float), ("f64", C.c_double)):
        fn = getattr(lib, f"oracle_aggregate_{name}")
        fn.argtypes = [C.POINTER(ct), C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(ct)]
        fn.restype = C.c_int
    lib.oracle_grid_new.argtypes = [C.c_int, C.c_int, i32p]
    lib.oracle_grid_new.restype = vp
    lib.oracle_grid_free.argtypes = [vp]
    lib.oracle_grid_free.restype = None
    lib.oracle_grid_insert.argtypes = [vp, C.c_uint32, i32p]
    lib.oracle_grid_insert.restype = C.c_int
    lib.oracle_grid_remove.argtypes = [vp, C.c_uint32]
    lib.oracle_grid_remove.restype = None
    lib.oracle_grid_query.argtypes = [vp, C.c_int, i32p, u32p, C.c_uint64]
    lib.oracle_grid_query.restype = C.c_uint64
    lib.oracle_grid_num_entities.argtypes = [vp]
    lib.oracle_grid_num_entities.restype = C.c_uint64
    lib.oracle_grid_in_bounds.argtypes = [vp, i32p]
    lib.oracle_grid_in_bounds.restype = C.c_int
    lib.oracle_neighbor_offsets.argtypes = [C.c_int, C.c_int, i32p]
    lib.oracle_neighbor_offsets.restype = C.c_uint64
    _lib = lib
    return lib
