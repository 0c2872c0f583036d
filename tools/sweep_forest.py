import glob, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = """
import sys; sys.path.insert(0, %r)
from tools.quick_bench import forest
forest(4096, 4096, 320); forest(16384, 16384, 20)
""" % root
for lib in sorted(glob.glob(os.path.join(root, "incerto_b200/lib/libincerto_b200*.so"))):
    env = dict(os.environ, INCERTO_B200_LIB=lib)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("==", os.path.basename(lib))
    print("\n".join(l for l in out.stdout.splitlines() if l.startswith("forest")) or out.stderr[-500:])
