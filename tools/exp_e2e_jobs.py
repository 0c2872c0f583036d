"""Per-job phase times of bench.py's N = 1 end-to-end job, run back to back (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import incerto_b200 as ib
from incerto_b200 import _ffi as F

W = H = 4096
pos, st = bench.host_forest(np, F, W, H, 0, W, [[7, 9], [2000, 2000], [2047, 0]])
for j in range(12):
    r = bench.e2e_job(ib, F, W, H, 0, pos, st, 500, None, None, W * H + 3)
    print(j, {k: round(v * (1 if k.endswith("_ms") else 1e3), 2) for k, v in r.items()}, flush=True)
