"""CPU-only checks of the air_traffic_3d oracle (examples/air_traffic_3d.rs): RNG-free known answers derived from
the reference's own semantics (neighbors_of excludes the centre cell, spatial_grid.rs:331-350; conflict iff
length_squared <= 1, air_traffic_3d.rs:140-141) and determinism of the Philox-keyed run."""
import numpy as np

from conftest import SEED
from oracle.air import AirOracle


def test_conflict_known_answers():
    xyz = np.array([[5, 5, 5], [5, 5, 6], [1, 1, 1], [1, 1, 1], [9, 9, 2], [10, 10, 2], [3, 7, 0], [4, 7, 0], [2, 7, 0],
                    [20, 20, 10], [19, 20, 10]], dtype=np.int32)
    o = AirOracle(SEED, 0, 0, systems=1).spawn_explicit(xyz, xyz[:, 2].copy()).run(1)
    assert o.series().tolist() == [4]


def test_altitude_converges_and_positions_clamped():
    o = AirOracle(SEED, 0, 500, 20, 10).spawn()
    xyz0, tgt = o.state()
    assert xyz0[:, :2].min() >= 0 and xyz0[:, :2].max() < 20 and xyz0[:, 2].max() < 10 and tgt.max() < 10
    o.run(12)
    xyz, _ = o.state()
    assert np.array_equal(xyz[:, 2], tgt)                      # one level per step towards the target (:98-105)
    assert xyz[:, :2].min() >= 0 and xyz[:, :2].max() <= 19   # clamp(0, AIRSPACE_SIZE - 1) (:113-114)
    assert np.abs(xyz[:, :2] - xyz0[:, :2]).max() <= 12


def test_out_of_bounds_panics_like_the_reference():
    import pytest
    o = AirOracle(SEED, 0, 0).spawn_explicit(np.array([[21, 0, 0]], np.int32), [0])
    with pytest.raises(IndexError):
        o.run(1)


def test_replicas_differ_and_runs_repeat():
    a = AirOracle(SEED, 0, 200).spawn().run(20).state()[0]
    b = AirOracle(SEED, 0, 200).spawn().run(20).state()[0]
    c = AirOracle(SEED, 1, 200).spawn().run(20).state()[0]
    assert np.array_equal(a, b) and not np.array_equal(a, c)
