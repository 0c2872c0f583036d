"""GPU: the CUDA path through the C ABI reproduces the committed golden vectors (no oracle in the loop)."""
import os

import numpy as np
import pytest

from golden_cases import CASES, compare_engine_to_golden, run_engine

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "workloads.npz"))


@pytest.mark.parametrize("name", [c for c in CASES if c != "normal_draws"])
def test_engine_reproduces_golden(name, engine_lib):
    gold = {k.split("/", 1)[1]: GOLD[k] for k in GOLD.files if k.startswith(name + "/")}
    compare_engine_to_golden(name, run_engine(name), gold)
