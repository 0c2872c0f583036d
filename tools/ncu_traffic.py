"""DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every step kernel, from the `ncu --set full`
captures of tools/ncu_all.sh:   python tools/ncu_traffic.py gpurun_out/r02_*.ncu-rep > profiles/r02_dram_traffic.json
Keys: bench.py workload name -> the engine's kernel-stat name -> bytes per launch (mean over the captured launches).
Runs without a GPU (ncu -i)."""
import csv
import json
import os
import re
import subprocess
import sys

WORKLOAD = {"coin": "coin_toss", "forest": "forest_fire_16384", "forest4096": "forest_fire_4096", "forestbench": "forest_fire_4096",
            "pandemic": "pandemic", "spatial": "pandemic_spatial", "traders": "traders", "air": "air_traffic_3d"}
KERNEL = [(r"coin_toss", "coin_toss_step"), (r"coin_odds", "coin_odds"), (r"forest_step", "forest_step"),
          (r"pandemic_move", "pandemic_move"), (r"pandemic_interact", "pandemic_interact"),
          (r"spatial_prepare", "spatial_prepare"), (r"spatial_interact", "spatial_interact"),
          (r"traders_step", "traders_step"), (r"stocks_price", "stocks_price_change"),
          (r"air_candidates", "air_candidates"), (r"air_step", "air_step"), (r"air_bits", "air_bits"),
          (r"air_count", "air_count"), (r"air_mark", "air_mark")]


def num(x):
    return float(x.replace(",", "")) if x not in ("", "n/a") else 0.0


def unit_scale(u):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


out = {}
for rep in sys.argv[1:]:
    m = re.search(r"r\d+_([a-z0-9]+)\.ncu-rep$", os.path.basename(rep))
    if not m or m.group(1) not in WORKLOAD:
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    acc = {}
    for r in rows[2:]:
        name = next((stat for pat, stat in KERNEL if re.search(pat, r[ik])), None)
        if name is None:
            continue
        b = num(r[ir]) * unit_scale(units[ir]) + num(r[iw]) * unit_scale(units[iw])
        acc.setdefault(name, []).append(b)
    w = out.setdefault(WORKLOAD[m.group(1)], {})
    for k, v in acc.items():
        w[k] = sum(v) / len(v)
out["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the launches captured by `ncu --set full "
                "--clock-control none` (tools/ncu_all.sh); ncu's per-kernel cache control makes every captured launch start "
                "with a cold L2")
print(json.dumps(out, indent=1))
