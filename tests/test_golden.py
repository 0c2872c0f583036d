"""CPU: the oracle keeps reproducing the committed golden vectors (tests/golden/workloads.npz, made by
tests/golden/make_golden.py) bit for bit - integer state and, thanks to the round-to-nearest-only normal draw,
f64 state as well."""
import os

import numpy as np
import pytest

from golden_cases import CASES, run_oracle

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "workloads.npz"))


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(name, oracle_lib):
    got = run_oracle(name)
    keys = [k for k in GOLD.files if k.startswith(name + "/")]
    assert sorted(k.split("/", 1)[1] for k in keys) == sorted(got)
    for k in keys:
        a, b = got[k.split("/", 1)[1]], GOLD[k]
        assert a.shape == b.shape and a.dtype == b.dtype, k
        if a.dtype.kind == "f":
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), k
        else:
            assert np.array_equal(a, b), k


def test_golden_is_not_trivial():
    assert GOLD["forest_fire/series"][-1][2] > 0            # something burned down
    lot = GOLD["forest_fire_lottery/series"]
    assert lot[-1][2] > 0 and 150 < int(lot[0][3]) - int(lot[-1][3]) < 350   # Empty cells regrew through the block lottery
    assert 200 < (~GOLD["pandemic_two_level/alive"]).sum() < 1500 and GOLD["pandemic_two_level/counts"][1] > 100   # deaths through the two-level fate draw
    assert (~GOLD["pandemic/alive"]).any() and (~GOLD["pandemic_spatial/alive"]).any()   # despawns happened
    assert GOLD["traders/delisted"].any() or GOLD["traders/price"].min() < 5.0
    assert GOLD["air_traffic_3d/conflicts"].max() > 0
    assert (GOLD["pandemic_spatial/quarantined"] > 0).any()
