"""Time every built variant of the engine (incerto_b200/lib/libincerto_b200.*.so) with tools/cold_forest.py."""
import glob, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sizes = sys.argv[1:] or ["4096x4096", "16384x16384"]
for lib in sorted(glob.glob(os.path.join(root, "incerto_b200/lib/libincerto_b200*.so"))):
    if lib.endswith(".trace.so"):
        continue
    env = dict(os.environ, INCERTO_B200_LIB=lib)
    out = subprocess.run([sys.executable, os.path.join(root, "tools/cold_forest.py"), *sizes], env=env, capture_output=True, text=True)
    res = " | ".join(l.replace("forest ", "") for l in out.stdout.splitlines() if l.startswith("forest")) or out.stderr[-300:]
    print("==", os.path.basename(lib), "|", res, flush=True)
