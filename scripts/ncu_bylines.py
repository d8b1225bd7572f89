#!/usr/bin/env python
"""Per-source-line warp instructions PER WARP TILE (128 hits) of an ncu report with source: python scripts/ncu_bylines.py rep hits [kernel regex] [min]"""
import csv, subprocess, sys
rep = sys.argv[1]; hits = float(sys.argv[2]); kern = sys.argv[3] if len(sys.argv) > 3 else "k_batch"; mn = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname = func = None; hdr = None; agg = {}; first = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name":
        func = r[1]
        if first is None: first = func
        continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or func != first: continue
    if r[2] != "-": continue
    try:
        ie = int(r[hdr.index("Instructions Executed")]); smp = int(r[hdr.index("# Samples")]); te = int(r[hdr.index("Thread Instructions Executed")])
    except ValueError: continue
    a = agg.setdefault((fname, int(r[0])), [0, 0, 0, r[1].strip()[:100]]); a[0] += ie; a[1] += smp; a[2] += te
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values()); nt = hits / 128
print("total warp inst %d = %.1f per tile; samples %d" % (tot, tot / nt, ts))
for (f, ln), a in sorted(agg.items()):
    if a[0] / nt >= mn or 100.0 * a[1] / ts >= 1.0:
        print("%-22s %4d %7.1f %5.1f%% thr %4.1f  %s" % (f, ln, a[0] / nt, 100.0 * a[1] / ts, a[2] / max(a[0], 1), a[3]))
