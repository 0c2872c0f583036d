"""Build tuning variants of the engine (forest kernel: min blocks per SM / warps per CTA / columns per thread)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
VARIANTS = [(10, 4, 4), (12, 2, 8), (6, 2, 16), (4, 2, 24)]
if __name__ == "__main__":
    for mb, ty, cols in VARIANTS:
        code = (f"from incerto_b200 import _build; print(_build.build(extra_flags=['-DINCERTO_FOREST_MIN_BLOCKS={mb}', "
                f"'-DINCERTO_FOREST_TY={ty}', '-DINCERTO_FOREST_COLS={cols}'], variant='mb{mb}_ty{ty}_c{cols}'))")
        subprocess.run([sys.executable, "-c", code], check=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
