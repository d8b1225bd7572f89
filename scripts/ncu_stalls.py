#!/usr/bin/env python
"""Top source lines by warp-stall samples, with the dominant stall reasons: python scripts/ncu_stalls.py report.ncu-rep [kernel regex] [top N]"""
import csv, subprocess, sys
rep = sys.argv[1]; kern = sys.argv[2] if len(sys.argv) > 2 else "k_batch"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname = func = first = None; hdr = None; agg = {}; tot = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name":
        func = r[1]
        if first is None: first = func
        continue
    if r[0] == "Line No": hdr = r; stall_cols = [(i, c) for i, c in enumerate(r) if c.startswith("stall_") and "Not Issued" not in c]; continue
    if hdr is None or func != first or r[2] != "-": continue
    try: smp = int(r[hdr.index("# Samples")])
    except ValueError: continue
    key = (fname, r[0]); a = agg.setdefault(key, [0, {}, r[1].strip()[:90]]); a[0] += smp
    for i, c in stall_cols:
        try: v = int(r[i])
        except ValueError: v = 0
        a[1][c] = a[1].get(c, 0) + v; tot[c] = tot.get(c, 0) + v
T = sum(a[0] for a in agg.values())
print("samples", T, " by reason:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(T, 1)) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    rs = ", ".join("%s %d" % (k[6:], v) for k, v in sorted(a[1].items(), key=lambda kv: -kv[1])[:3] if v)
    print("%6d %5.1f%%  %s:%s  %s   [%s]" % (a[0], 100.0 * a[0] / max(T, 1), f, ln, a[2], rs))
