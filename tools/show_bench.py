"""Condensed view of a bench.py JSON line:  python tools/show_bench.py gpurun_out/r02_bench_x.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]
print(f"headline: {d['config']['workload']}: {d['ms_per_step'] * 1e3:.2f} us/step, {d['value'] / 1e9:.1f} G/s, frac {r['frac']:.3f}; "
      f"L2-resident {d['value_l2_resident'] / 1e9:.1f} G/s; launches {d['gpu_launches']}")
for k in ("same_slab_1gpu_us", "same_slab_1gpu_l2_resident_us", "fixed_cost_us", "fixed_cost_l2_resident_us", "partition_parity", "replica_parity", "global_aggregate_parity", "air_partition_parity"):
    if k in d:
        print(f"  {k}: {d[k]}")
if "replica_stream" in d:
    q = d["replica_stream"]
    print(f"replica_stream: {q['replicas_in_flight']} x 4096^2 in flight: {q['us_per_replica_step']:.2f} us per replica-step, {q['achieved_GBps']:.0f} GB/s, frac {q['frac']:.3f}")
if "e2e" in d:
    e = d["e2e"]
    print(f"e2e: best {e['value'] / 1e9:.1f} G/s ({e['seconds'] * 1e3:.2f} ms), median {e['median'] / 1e9:.1f} G/s; {e['breakdown_best_job_ms']}")
for name, w in d.get("workloads", {}).items():
    cold = w.get("cold_l2")
    print(f"{name:18s} {w['us_per_step']:8.1f} us/step {w['updates_per_s'] / 1e9:8.2f} G/s  {w['algorithmic_bytes_per_update']:6.1f} B/upd  frac {w['frac']:.3f}"
          + (f"  cold {cold['us_per_step']:.1f} us frac {cold['frac']:.3f}" if cold else ""))
    for kn, k in w["kernels"].items():
        print(f"      {kn:22s} {k['us_per_launch']:8.1f} us x {k['launches_per_step']:.2f}  {k['algorithmic_MB_per_launch']:8.1f} MB  frac {k['frac'] if k['frac'] is None else round(k['frac'], 3)}"
              + (f"  dram {k['dram_bytes_per_launch_ncu'] / 1e6:.0f} MB" if "dram_bytes_per_launch_ncu" in k else ""))
for name, w in d.get("replicas", {}).items():
    print(f"replicas {name}: {w['updates_per_s'] / 1e9:.2f} G/s, {w['replicas']} replicas, {w['us_per_replica_step']:.1f} us per replica-step, frac/GPU {w['frac_per_gpu']:.3f}")
if "air_one_world" in d:
    q = d["air_one_world"]
    print(f"air_one_world: {q['us_per_step']:.1f} us/step, {q['updates_per_s'] / 1e9:.2f} G/s, conflicts {q['conflicts_last_step']}, owned {q['aircraft_owned_by_all_ranks']}")
if "cpu_baseline" in d:
    print("cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"])
