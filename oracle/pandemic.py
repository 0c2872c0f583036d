"""ORACLE — test infrastructure, not product code.  numpy wrapper over the C++ pandemic oracle."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .lib import load


class PandemicOracle:
    def __init__(self, seed, replica, population, grid_size, p_start=0.05, p_infect=0.02, p_recover=0.001, p_die=0.0005,
                 systems=15, nested_loop=True):
        self.lib = load()
        self.h = self.lib.oracle_pandemic_new(seed, replica, population, grid_size, p_start, p_infect, p_recover, p_die,
                                              systems, int(nested_loop))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.oracle_pandemic_free(self.h)
            self.h = None

    def spawn(self):
        self.lib.oracle_pandemic_spawn(self.h)
        return self

    def spawn_explicit(self, xy, infected):
        xy = np.ascontiguousarray(xy, dtype=np.uint16)
        inf = np.ascontiguousarray(infected, dtype=np.uint8)
        self.lib.oracle_pandemic_spawn_explicit(self.h, len(inf), xy.ctypes.data_as(C.POINTER(C.c_uint16)),
                                                inf.ctypes.data_as(C.POINTER(C.c_uint8)))
        return self

    def run(self, steps):
        self.lib.oracle_pandemic_run(self.h, steps)
        return self

    def state(self):
        n = self.lib.oracle_pandemic_num_slots(self.h)
        xy = np.zeros((n, 2), np.uint16)
        alive = np.zeros(n, np.uint8)
        inf = np.zeros(n, np.uint8)
        self.lib.oracle_pandemic_state(self.h, xy.ctypes.data_as(C.POINTER(C.c_uint16)),
                                       alive.ctypes.data_as(C.POINTER(C.c_uint8)), inf.ctypes.data_as(C.POINTER(C.c_uint8)))
        return xy, alive.astype(bool), inf.astype(bool)

    def counts(self):
        out = np.zeros(3, np.uint64)
        self.lib.oracle_pandemic_counts(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return tuple(int(v) for v in out)
