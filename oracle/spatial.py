"""ORACLE — test infrastructure, not product code.  numpy wrapper over the C++ pandemic_spatial oracle."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from .lib import load

_I32P = C.POINTER(C.c_int32)
_U8P = C.POINTER(C.c_uint8)
_U32P = C.POINTER(C.c_uint32)
_U64P = C.POINTER(C.c_uint64)

HEALTHY, INFECTIOUS, RECOVERED = 0, 0x80, 0x81


@dataclass
class SpatialConfig:
    """examples/pandemic_spatial.rs:28-58"""
    population: int = 2000
    grid_size: int = 40
    chance_start_infected: float = 0.02
    infection_radius: int = 2
    chance_infect_1: float = 0.15
    chance_infect_2: float = 0.05
    chance_recover: float = 0.03
    chance_die: float = 0.001
    incubation_period: int = 5
    crowding_threshold: int = 8
    social_compliance: float = 0.7
    quarantine_duration: int = 14
    zone_x: int | None = None      # GRID_SIZE / 2
    zone_y: int | None = None
    zone_radius: int = 8
    history_clear_above: int = 20
    chance_attempt_move: float = 0.5
    systems: int = 0x7f            # bit k: system kind 40 + k

    def zone(self):
        return (self.grid_size // 2 if self.zone_x is None else self.zone_x,
                self.grid_size // 2 if self.zone_y is None else self.zone_y)


class SpatialOracle:
    def __init__(self, seed, replica, cfg: SpatialConfig):
        self.lib = load()
        self.cfg = cfg
        zx, zy = cfg.zone()
        prob = (C.c_double * 7)(cfg.chance_start_infected, cfg.chance_infect_1, cfg.chance_infect_2, cfg.chance_recover,
                                cfg.chance_die, cfg.social_compliance, cfg.chance_attempt_move)
        ints = (C.c_int32 * 9)(cfg.infection_radius, cfg.incubation_period, cfg.crowding_threshold,
                               cfg.quarantine_duration, zx, zy, cfg.zone_radius, cfg.history_clear_above, 0)
        self.h = self.lib.oracle_spatial_new(seed, replica, cfg.population, cfg.grid_size, prob, ints, cfg.systems)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.oracle_spatial_free(self.h)
            self.h = None

    def spawn(self):
        self.lib.oracle_spatial_spawn(self.h)
        return self

    def spawn_explicit(self, xy, disease, social):
        xy = np.ascontiguousarray(xy, dtype=np.int32)
        d = np.ascontiguousarray(disease, dtype=np.uint8)
        s = np.ascontiguousarray(social, dtype=np.uint8)
        self.lib.oracle_spatial_spawn_explicit(self.h, len(d), xy.ctypes.data_as(_I32P), d.ctypes.data_as(_U8P),
                                               s.ctypes.data_as(_U8P))
        return self

    def run(self, steps):
        if self.lib.oracle_spatial_run(self.h, steps):
            raise IndexError("entity outside spatial grid bounds")
        return self

    def state(self):
        n = self.lib.oracle_spatial_num_slots(self.h)
        xy = np.zeros((n, 2), np.int32)
        alive = np.zeros(n, np.uint8)
        disease = np.zeros(n, np.uint8)
        social = np.zeros(n, np.uint8)
        quar = np.zeros(n, np.uint32)
        hist = np.zeros((n, 21), np.uint32)
        self.lib.oracle_spatial_state(self.h, xy.ctypes.data_as(_I32P), alive.ctypes.data_as(_U8P),
                                      disease.ctypes.data_as(_U8P), social.ctypes.data_as(_U8P),
                                      quar.ctypes.data_as(_U32P), hist.ctypes.data_as(_U32P))
        return dict(xy=xy, alive=alive.astype(bool), disease=disease, social=social.astype(bool), quarantined=quar,
                    history=hist)

    def stats(self):
        out = np.zeros(6, np.uint64)
        self.lib.oracle_spatial_stats(self.h, out.ctypes.data_as(_U64P))
        return tuple(int(v) for v in out)

    def series(self):
        n = self.lib.oracle_spatial_series(self.h, None, 0)
        out = np.zeros((n, 6), np.uint64)
        self.lib.oracle_spatial_series(self.h, out.ctypes.data_as(_U64P), n)
        return out
