"""Generates tests/golden/workloads.npz: small seeded input/output vectors of every workload, produced by the CPU
oracle (oracle/).  The reference itself (Rust on Bevy) cannot run in this environment and none of its tests pins an
RNG-dependent result, so these vectors pin OUR draw contract: the oracle must keep reproducing them
(tests/test_golden.py, CPU) and the CUDA path must reproduce them without the oracle in the loop
(tests/test_gpu_golden.py).  Re-run only when the draw contract changes on purpose:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_cases import CASES, run_oracle  # noqa: E402

if __name__ == "__main__":
    out = {}
    for name in CASES:
        for k, v in run_oracle(name).items():
            out[f"{name}/{k}"] = v
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "workloads.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", len(out), "arrays")
