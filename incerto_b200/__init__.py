"""incerto_b200 — B200-native (sm_100a) execution engine for incerto's per-step simulation hot path.

The package holds the CUDA engine (csrc/, built into lib/libincerto_b200.so) and the host-side
mirror of the reference's SimulationBuilder / Simulation surface.  There is no CPU fallback: the
library must be built (`python -m incerto_b200._build`) and a CUDA device must be present.
"""
from . import _ffi
from .simulation import *  # noqa: F401,F403
from .simulation import (Aggregate, BuilderError, Component, DeviceSpawner, EngineError, SamplingError,  # noqa: F401
                         Simulation, SimulationBuilder, SimulationError, Spawner, System, TimeSeries, With, Without,
                         comm_unique_id, partition_range)

__version__ = "0.1.0"
