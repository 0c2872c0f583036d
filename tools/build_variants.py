"""Build tuning variants of the engine (forest kernel launch bounds / tile shape)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for mb, ty in ((3, 8), (5, 8), (6, 8), (4, 4), (8, 4), (6, 4), (10, 4), (2, 16)):
    code = (f"from incerto_b200 import _build; print(_build.build(extra_flags=['-DINCERTO_FOREST_MIN_BLOCKS={mb}', "
            f"'-DINCERTO_FOREST_TY={ty}'], variant='mb{mb}_ty{ty}'))")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
