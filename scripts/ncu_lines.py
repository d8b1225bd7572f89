#!/usr/bin/env python
"""Per-source-line summary of an ncu report (needs -lineinfo + --import-source on):
   python scripts/ncu_lines.py report.ncu-rep [kernel regex] [top N]"""
import csv, subprocess, sys
rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else "k_batch"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname = func = None
hdr = None
agg = {}
first_func = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name":
        func = r[1]
        if first_func is None: first_func = func
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or func != first_func:
        continue
    if r[2] != "-":   # SASS rows carry an address; the line rows carry '-'
        continue
    try:
        ie = int(r[hdr.index("Instructions Executed")]); te = int(r[hdr.index("Thread Instructions Executed")])
        smp = int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    key = (fname, r[0])
    a = agg.setdefault(key, [0, 0, 0, r[1].strip()[:100]])
    a[0] += ie; a[1] += te; a[2] += smp
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
print("kernel:", (first_func or "")[:120])
print("total warp instructions %d, samples %d" % (tot, tots))
print("%9s %6s %6s %5s  %s" % ("inst", "inst%", "smpl%", "thr", "line"))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%9d %5.1f%% %5.1f%% %5.1f  %s:%s  %s" % (a[0], 100.0 * a[0] / max(tot, 1), 100.0 * a[2] / max(tots, 1), a[1] / max(a[0], 1), f, ln, a[3]))

# ---- by line ranges (optional 4th argument: comma separated "name:lo-hi")
if len(sys.argv) > 4:
    print()
    for spec in sys.argv[4].split(","):
        name, rng = spec.split(":"); lo, hi = map(int, rng.split("-"))
        sel = [a for (f, ln), a in agg.items() if f.startswith("mma_device") and lo <= int(ln) <= hi]
        i = sum(a[0] for a in sel); t = sum(a[1] for a in sel); sm_ = sum(a[2] for a in sel)
        print("%-14s inst %9d (%5.1f%%)  samples %5.1f%%  avg threads %4.1f" % (name, i, 100.0 * i / max(tot, 1), 100.0 * sm_ / max(tots, 1), t / max(i, 1)))
    other = [a for (f, ln), a in agg.items() if not f.startswith("mma_device")]
    print("%-14s inst %9d (%5.1f%%)  samples %5.1f%%" % ("other files", sum(a[0] for a in other), 100.0 * sum(a[0] for a in other) / max(tot, 1), 100.0 * sum(a[2] for a in other) / max(tots, 1)))
