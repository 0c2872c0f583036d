"""ORACLE — test infrastructure, not product code.  numpy wrapper over the C++ air_traffic_3d oracle."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .lib import load

_I32P = C.POINTER(C.c_int32)


class AirOracle:
    def __init__(self, seed, replica, n, size=20, height=10, p_move=0.7, systems=3):
        self.lib = load()
        self.h = self.lib.oracle_air_new(seed, replica, n, size, height, p_move, systems)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.oracle_air_free(self.h)
            self.h = None

    def spawn(self):
        self.lib.oracle_air_spawn(self.h)
        return self

    def spawn_explicit(self, xyz, target):
        xyz = np.ascontiguousarray(xyz, dtype=np.int32)
        target = np.ascontiguousarray(target, dtype=np.int32)
        self.lib.oracle_air_spawn_explicit(self.h, len(target), xyz.ctypes.data_as(_I32P), target.ctypes.data_as(_I32P))
        return self

    def run(self, steps):
        rc = self.lib.oracle_air_run(self.h, steps)
        if rc:
            raise IndexError("entity outside spatial grid bounds")   # spatial_grid.rs:276-279 panics
        return self

    def state(self):
        n = self.lib.oracle_air_num(self.h)
        xyz = np.zeros((n, 3), np.int32)
        target = np.zeros(n, np.int32)
        self.lib.oracle_air_state(self.h, xyz.ctypes.data_as(_I32P), target.ctypes.data_as(_I32P))
        return xyz, target

    def series(self):
        n = self.lib.oracle_air_series(self.h, None, 0)
        out = np.zeros(n, np.uint64)
        self.lib.oracle_air_series(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64)), n)
        return out
