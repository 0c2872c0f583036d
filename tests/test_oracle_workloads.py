"""CPU-only, RNG-free known answers of the traders and pandemic_spatial oracles, derived from the reference's own
semantics (examples/traders.rs, examples/pandemic_spatial.rs) with every probability at 0 or 1."""
import numpy as np
import pytest

from conftest import SEED
from oracle.spatial import HEALTHY, INFECTIOUS, RECOVERED, SpatialConfig, SpatialOracle
from oracle.traders import TradersOracle, det_normal


# ------------------------------------------------------------------ traders
def test_traders_buy_until_cash_runs_out_then_sell_everything():
    """p_buy = 1, two shares at 5.0 each: ten purchases exhaust the 100.0 of cash, the eleventh is skipped
    (amount > cash, traders.rs:184-188); net worth stays 100 (:250-261); p_sell = 1 sells all of it back (:229-233)."""
    o = TradersOracle(SEED, 0, 3, 25, p_buy=1.0, p_sell=0.0, shares=(2, 2), systems=1 | 4).run(1)
    st = o.state()
    assert (st["cash"] == 0.0).all() and (st["net_worth"] == 100.0).all()
    assert (st["shares"][:, :10] == 2).all() and (st["shares"][:, 10:] == 0).all()
    o2 = TradersOracle(SEED, 0, 3, 25, p_buy=1.0, p_sell=1.0, shares=(2, 2), systems=1 | 2 | 4).run(1)
    st2 = o2.state()
    assert (st2["cash"] == 100.0).all() and (st2["shares"] == 0).all() and (st2["net_worth"] == 100.0).all()


def test_stocks_delist_at_zero_and_are_never_bought_again():
    """mean -1, std 0: 5 -> 4 -> ... -> 0; `price < f64::EPSILON` inserts StockDelisted (:157-161) and the buy query
    is Without<StockDelisted> (:167)."""
    o = TradersOracle(SEED, 0, 2, 4, p_buy=0.0, p_sell=0.0, mean=-1.0, std_dev=0.0).run(4)
    assert (o.state()["price"] == 1.0).all() and not o.state()["delisted"].any()
    o.run(1)
    assert (o.state()["price"] == 0.0).all() and o.state()["delisted"].all()
    o.run(3)
    assert (o.state()["price"] == 0.0).all()                         # Without<StockDelisted>: no further changes (:142)
    b = TradersOracle(SEED, 0, 2, 4, p_buy=1.0, p_sell=0.0, shares=(1, 1), mean=-5.0, std_dev=0.0).run(3)
    st = b.state()
    assert (st["shares"] == 1).all() and (st["cash"] == 80.0).all()   # bought at step 1 (price 5), nothing later
    assert (st["net_worth"] == 80.0).all()                            # delisted shares are worth 0 (:254-258)


def test_normal_draw_is_symmetric_and_finite_at_the_corners():
    assert det_normal([0, 0, 0, 0]) > 8.0                             # u1 = 2^-53, u2 = 0: the far right tail
    assert det_normal([0xFFFFFFFF] * 4) == pytest.approx(0.0, abs=1e-7)  # u1 = 1: r = 0
    z = np.array([det_normal(w) for w in np.random.default_rng(3).integers(0, 2**32, size=(4000, 4))])
    assert abs(z.mean()) < 0.06 and abs(z.std() - 1.0) < 0.05


# ---------------------------------------------------------- pandemic_spatial
def spatial(xy, disease, social=None, **kw):
    cfg = SpatialConfig(population=len(disease), **kw)
    return SpatialOracle(SEED, 0, cfg).spawn_explicit(np.array(xy), disease, social or [0] * len(disease))


def test_incubation_counts_down_to_infectious():
    o = spatial([[1, 1]], [3], grid_size=4, systems=1)               # Exposed{3} (:414-433)
    assert [int(o.run(1).state()["disease"][0]) for _ in range(4)] == [2, 1, INFECTIOUS, INFECTIOUS]


def test_transmission_reaches_manhattan_distance_two_only():
    xy = [[5, 5], [5, 6], [6, 6], [5, 8], [5, 5]]                     # source, d=1, d=2, d=3, same cell
    o = spatial(xy, [INFECTIOUS, HEALTHY, HEALTHY, HEALTHY, HEALTHY], grid_size=12, chance_infect_1=1.0,
                chance_infect_2=1.0, systems=2).run(1)
    assert o.state()["disease"].tolist() == [INFECTIOUS, 5, 5, HEALTHY, 5]   # Exposed{INCUBATION_PERIOD} (:510-517)


def test_recovery_and_death_are_exclusive_and_death_wins():
    o = spatial([[0, 0], [1, 1]], [INFECTIOUS, HEALTHY], grid_size=3, chance_die=0.0, chance_recover=1.0, systems=4).run(1)
    assert o.state()["disease"].tolist() == [RECOVERED, HEALTHY]
    o = spatial([[0, 0], [1, 1]], [INFECTIOUS, HEALTHY], grid_size=3, chance_die=1.0, chance_recover=1.0, systems=4).run(1)
    assert o.state()["alive"].tolist() == [False, True] and o.stats() == (1, 0, 0, 0, 1, 1)   # dead = INITIAL - alive (:134)


def test_contact_tracing_quarantines_cell_mates_for_fourteen_steps():
    """Two people sharing a cell remember it and quarantine each other at the first flush (:565-608); the counter goes
    14 -> 1 and the component is removed one step later (:611-625); a person alone is never quarantined."""
    o = spatial([[2, 2], [2, 2], [7, 7]], [HEALTHY] * 3, grid_size=10, systems=8 | 16 | 32)
    seen = []
    for _ in range(17):
        o.run(1)
        seen.append(o.state()["quarantined"].tolist())
    assert seen[0] == [14, 14, 0] and seen[1] == [13, 13, 0] and seen[13] == [1, 1, 0]
    assert seen[14] == [0, 0, 0]                                      # removed at the flush of step 15
    assert seen[15] == [14, 14, 0]                                    # free again during step 16: traced again
    assert o.state()["history"][:, 0].tolist() == [1, 1, 1]           # nobody moved: one remembered cell each


def test_aggregate_series_skips_steps_without_people():
    o = spatial([[0, 0]], [INFECTIOUS], grid_size=2, chance_die=1.0, systems=4).run(5)
    assert len(o.series()) == 0                                       # `if !component_values.is_empty()` (time_series.rs:79)
