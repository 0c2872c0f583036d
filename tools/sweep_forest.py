"""Time every built variant of the engine on the bench.py configuration (forest 4096^2, L2 flushed) and at 16384^2."""
import glob, json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = """
import sys; sys.path.insert(0, %r)
from tools.quick_bench import forest
forest(16384, 16384, 20)
""" % root
for lib in sorted(glob.glob(os.path.join(root, "incerto_b200/lib/libincerto_b200*.so"))):
    env = dict(os.environ, INCERTO_B200_LIB=lib)
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "100", "--no-cpu-baseline", "--no-e2e"],
                         env=env, capture_output=True, text=True)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
        cold = f"4096^2 cold {j['ms_per_step'] * 1e3:.2f} us/step (frac {j['roofline']['frac']:.3f}), L2-resident {16777216 / j['value_l2_resident'] * 1e6:.2f} us"
    except Exception:
        cold = "bench failed: " + out.stderr[-300:]
    out2 = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    big = " | ".join(l.split(":", 1)[1].split(",")[0].strip() for l in out2.stdout.splitlines() if l.startswith("forest")) or out2.stderr[-200:]
    print("==", os.path.basename(lib), "|", cold, "| 16384^2", big, flush=True)
