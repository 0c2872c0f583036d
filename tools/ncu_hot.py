"""Top SASS instructions of each kernel in a .ncu-rep by stall samples / executed count (source page)."""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kernel, hdr, rows = None, None, []


def flush():
    if not rows:
        return
    tot = sum(r[0] for r in rows) or 1
    toti = sum(r[1] for r in rows)
    print(f"=== {kernel}: samples {tot}, warp-instr {toti}")
    for s, n, thr, src in sorted(rows, reverse=True)[:top]:
        print(f"  {100 * s / tot:5.1f}%  exec {n:>10d}  avg-thr {thr:>5s}  {src[:110]}")


for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "Kernel Name":
        flush()
        kernel, hdr, rows = r[1][:70], None, []
    elif r[0] == "Address":
        hdr = r
    elif hdr and len(r) == len(hdr):
        try:
            rows.append((int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")]),
                         r[hdr.index("Avg. Threads Executed")], r[hdr.index("Source")].strip()))
        except ValueError:
            pass
flush()
