"""Print the headline ncu metrics of a .ncu-rep (run here, no GPU needed)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__cycles_active.avg', 'smsp__cycles_active.avg']
for r in rows[2:]:
    print('--- kernel', r[hdr.index('Kernel Name')][:60], 'grid', r[hdr.index('Grid Size')], 'block', r[hdr.index('Block Size')])
    for w in want:
        if w in hdr:
            print(f'  {w:68s} {r[hdr.index(w)]} {units[hdr.index(w)]}')
    stall = [h for h in hdr if 'smsp__average_warp' in h and 'issue_stalled' in h and h.endswith('.ratio')]
    vals = sorted(((float(r[hdr.index(h)].replace(',', '')), h) for h in stall if r[hdr.index(h)] not in ('', 'n/a')), reverse=True)
    for v, h in vals[:7]:
        print(f'  {v:8.3f} {h.replace("smsp__average_warps_issue_stalled_", "stall:").replace("_per_issue_active.ratio", "")}')
