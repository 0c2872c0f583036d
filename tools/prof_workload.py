"""Run one workload of tools/bench_workloads.py for a few steps (to be wrapped by ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_workloads as bw


def short_report(name, sim, n, steps, bytes_per_update, warm=5, prof_steps=10):
    sim.run(int(os.environ.get("PROF_WARM", "12")))
    sim.run(int(os.environ.get("PROF_STEPS", "2")))
    print(name, "done", flush=True)


bw.report = short_report
if "--scale" in sys.argv:
    bw.SCALE = float(sys.argv[sys.argv.index("--scale") + 1])
getattr(bw, sys.argv[1])()
