"""ORACLE — test infrastructure, not product code.  Thin numpy wrapper over the C++ forest-fire oracle."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .lib import load


class ForestOracle:
    def __init__(self, seed, replica, width, height, density=0.7, p_spread=0.6, burn=3, p_regrow=0.001, fires=3,
                 systems=7):
        self.lib = load()
        self.h = self.lib.oracle_forest_new(seed, replica, width, height, density, p_spread, burn, p_regrow, fires,
                                            systems)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.oracle_forest_free(self.h)
            self.h = None

    def spawn(self):
        self.lib.oracle_forest_spawn(self.h)
        return self

    def spawn_explicit(self, xy: np.ndarray, states: np.ndarray):
        xy = np.ascontiguousarray(xy, dtype=np.int32)
        st = np.ascontiguousarray(states, dtype=np.uint8)
        self.lib.oracle_forest_spawn_explicit(self.h, len(st), xy.ctypes.data_as(C.POINTER(C.c_int32)),
                                              st.ctypes.data_as(C.POINTER(C.c_uint8)))
        return self

    def run(self, steps):
        self.lib.oracle_forest_run(self.h, steps)
        return self

    def cells(self):
        n = self.lib.oracle_forest_num_entities(self.h)
        st = np.zeros(n, np.uint8)
        xy = np.zeros((n, 2), np.int32)
        self.lib.oracle_forest_cells(self.h, st.ctypes.data_as(C.POINTER(C.c_uint8)),
                                     xy.ctypes.data_as(C.POINTER(C.c_int32)))
        return st, xy

    def stats(self):
        out = np.zeros(5, np.uint64)
        self.lib.oracle_forest_stats(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return out

    def series(self):
        n = self.lib.oracle_forest_series_len(self.h)
        out = np.zeros((n, 5), np.uint64)
        if n:
            self.lib.oracle_forest_series(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return out


def coin_toss_run(seed, replica, n, odds, first_step, steps, heads=None, tosses=None):
    lib = load()
    h = np.zeros(n, np.uint64) if heads is None else np.ascontiguousarray(heads, dtype=np.uint64).copy()
    t = np.zeros(n, np.uint64) if tosses is None else np.ascontiguousarray(tosses, dtype=np.uint64).copy()
    agg = lib.oracle_coin_toss_run(seed, replica, n, odds, first_step, steps, h.ctypes.data_as(C.POINTER(C.c_uint64)),
                                   t.ctypes.data_as(C.POINTER(C.c_uint64)))
    return h, t, agg
