"""coin_toss (examples/coin_toss.rs) on the GPU vs the CPU oracle: bit-exact heads, f64 aggregate."""
import numpy as np
import pytest

import incerto_b200 as ib
from incerto_b200 import _ffi as F
from conftest import SEED
from oracle.forest import coin_toss_run

pytestmark = pytest.mark.gpu


def _build(n, odds=0.5, dtype=F.U32, replica=0, device_spawn=False, series=None):
    b = ib.SimulationBuilder(seed=SEED, replica=replica)
    heads = b.component("CoinTosser.num_heads", dtype)
    tosses = b.component("CoinTosser.num_tosses", dtype)
    if device_spawn:
        b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_COIN_TOSSERS, F.SpawnCoinTossersParams(heads.handle, tosses.handle, n)))
    else:
        b.add_entity_spawner(lambda s: s.spawn_batch(n, {heads: np.zeros(n), tosses: np.zeros(n)}))
    b.add_systems(ib.CoinToss(heads, tosses, odds))
    if series:
        b.record_aggregate_time_series(ib.CoinOdds(heads, tosses), series)
    return b.build(), heads, tosses


def test_reference_config_as_is():
    """examples/coin_toss.rs:18-20: 100 entities, 10 000 steps, p = 0.5 (BASELINE config 0)."""
    sim, heads, tosses = _build(100)
    sim.run(10_000)
    h, t, agg = coin_toss_run(SEED, 0, 100, 0.5, 1, 10_000)
    assert np.array_equal(sim.read_component(heads), h)
    assert np.array_equal(sim.read_component(tosses), t)
    assert sim.sample_aggregate(ib.CoinOdds(heads, tosses)) == pytest.approx(agg, rel=1e-12)
    assert 0.49 < agg < 0.51


@pytest.mark.parametrize("n", [1, 3, 4, 5, 127, 1000, 4099, 65_537])
@pytest.mark.parametrize("dtype", [F.U32, F.U64])
def test_parity_ragged_sizes(n, dtype):
    sim, heads, tosses = _build(n, odds=0.3, dtype=dtype, replica=3)
    sim.run(37)
    h, t, agg = coin_toss_run(SEED, 3, n, 0.3, 1, 37)
    assert np.array_equal(sim.read_component(heads).astype(np.uint64), h)
    assert np.array_equal(sim.read_component(tosses).astype(np.uint64), t)
    assert sim.sample_aggregate(ib.CoinOdds(heads, tosses)) == pytest.approx(agg, rel=1e-12)


def test_single_steps_equal_fused_steps_and_continue():
    """run(n) may fuse steps in registers; one-at-a-time stepping must give the same bits (and run continues, test_counter.rs:57-61)."""
    n = 10_007
    a, ha, _ = _build(n, device_spawn=True)
    bsim, hb, _ = _build(n)
    a.run(64)
    for _ in range(64):
        bsim.run(1)
    assert np.array_equal(a.read_component(ha), bsim.read_component(hb))
    a.run(11)
    h, _, _ = coin_toss_run(SEED, 0, n, 0.5, 1, 75)
    assert np.array_equal(a.read_component(ha), h)
    assert a.step_number() == 76


def test_series_cadence_and_values():
    n = 1000
    sim, heads, tosses = _build(n, series=7)
    sim.run(30)
    ts = sim.get_aggregate_time_series(ib.CoinOdds(heads, tosses))
    assert ts.time() == [7, 14, 21, 28]
    for step, v in ts.enumerate():
        _, _, agg = coin_toss_run(SEED, 0, n, 0.5, 1, step)
        assert v == pytest.approx(agg, rel=1e-12)


def test_replicas_differ_and_reset_reproduces():
    sim, heads, _ = _build(5000, replica=1)
    sim.run(20)
    r1 = sim.read_component(heads).copy()
    sim.reset(2)
    sim.run(20)
    r2 = sim.read_component(heads).copy()
    sim.reset(1)
    sim.run(20)
    assert np.array_equal(sim.read_component(heads), r1) and not np.array_equal(r1, r2)
    h2, _, _ = coin_toss_run(SEED, 2, 5000, 0.5, 1, 20)
    assert np.array_equal(r2, h2)


def test_large_population_properties():
    """2^24 entities x 64 steps: size-independent checks (tosses == steps everywhere; heads ~ Binomial)."""
    n = 1 << 24
    sim, heads, tosses = _build(n, device_spawn=True)
    sim.run(64)
    assert sim.sample_aggregate(ib.Minimum(tosses)) == 64 and sim.sample_aggregate(ib.Maximum(tosses)) == 64
    total = sim.sample_aggregate(ib.Sum(heads))
    assert abs(total / (n * 64) - 0.5) < 1e-3
    # spot-check a window against the oracle: entity uid -> quad uid>>2 is independent of the population size
    h = sim.read_component(heads)
    ho, _, _ = coin_toss_run(SEED, 0, 4096, 0.5, 1, 64)
    assert np.array_equal(h[:4096], ho)
    assert sim.sample_aggregate(ib.CoinOdds(heads, tosses)) == pytest.approx(total / (n * 64), rel=1e-9)
