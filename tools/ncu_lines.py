"""Instructions executed / stall samples per CUDA source line of the first kernel in a .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kern = ["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []   # optional: kernel name pattern
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", *kern], capture_output=True, text=True).stdout
rows, fname, func, hdr, first_func = [], None, None, None, None
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        func = r[1]
        first_func = first_func or func
    elif r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and func == first_func and r[2] == "-":
        try:
            rows.append((int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")]), fname, r[0], r[1].strip()))
        except ValueError:
            pass
# the same kernel may appear once per captured launch: keep one copy per (file, line)
seen, uniq = set(), []
for x in rows:
    if (x[2], x[3]) not in seen:
        seen.add((x[2], x[3]))
        uniq.append(x)
tot_i = sum(x[0] for x in uniq) or 1
tot_s = sum(x[1] for x in uniq) or 1
print(f"{first_func[:80]}: warp-instr {tot_i}, samples {tot_s}")
for n, smp, f, ln, src in sorted(uniq, reverse=True)[:top]:
    print(f"  {100 * n / tot_i:5.1f}% instr  {100 * smp / tot_s:5.1f}% samples  {f}:{ln:>4s}  {src[:100]}")
