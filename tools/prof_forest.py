"""Short forest-fire run for ncu (W H steps)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_bench import forest

W, H, steps = (int(x) for x in sys.argv[1:4])
forest(W, H, steps)
