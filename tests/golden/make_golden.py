#!/usr/bin/env python
"""Regenerates the golden fixtures from the reference itself (run in the build container only).

Needs /root/reference (the shipped test_dataset.bam / test_dataset.gtf / configHS38.txt) and the
binaries built by oracle/build_ref.sh.  Writes
  tests/golden/chrY.npz            packed feature + hit buffers decoded by OUR host front-end
                                   (features verified line by line against `mmannot_dump`)
  tests/golden/chrY_expected.json  tables + statistics printed by the REFERENCE for the argument
                                   matrix of SURVEY.md section 4 (as-shipped binary for -s U,
                                   setFlags-repaired binary for -s F / -s R)
"""
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mmannot_b200 import host  # noqa: E402
from oracle import pyoracle  # noqa: E402

REF = os.environ.get("MMANNOT_REFERENCE", "/root/reference")
GTF, BAM, CFG = REF + "/test_dataset.gtf", REF + "/test_dataset.bam", REF + "/configHS38.txt"

MATRIX = [  # (name, binary kind, extra args)
    ("U", "asis", ["-s", "U"]),
    ("U_unique", "asis", ["-s", "U", "-y", "unique"]),
    ("U_random", "asis", ["-s", "U", "-y", "random"]),
    ("U_ratio", "asis", ["-s", "U", "-y", "ratio"]),
    ("U_l1", "asis", ["-s", "U", "-l", "1"]),
    ("U_l10", "asis", ["-s", "U", "-l", "10"]),
    ("U_l0.5", "asis", ["-s", "U", "-l", "0.5"]),
    ("U_l0.9", "asis", ["-s", "U", "-l", "0.9"]),
    ("U_d5000_D200", "asis", ["-s", "U", "-d", "5000", "-D", "200"]),
    ("U_e50", "asis", ["-s", "U", "-e", "50"]),
    ("U_e50_m", "asis", ["-s", "U", "-e", "50", "-m", "@M"]),
    ("F", "fixed", ["-s", "F"]),
    ("R", "fixed", ["-s", "R"]),
    ("F_l1", "fixed", ["-s", "F", "-l", "1"]),
    ("R_ratio", "fixed", ["-s", "R", "-y", "ratio"]),
]


def main():
    subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")])
    cfg = host.Config(CFG)
    ann = host.Annotation(cfg, GTF)
    # feature order check against the reference's own sorted interval list
    rc, out, err = pyoracle.run_reference(["-a", GTF, "-r", BAM, "-c", CFG, "-s", "U", "-o", os.devnull], kind="dump")
    ref = []
    for line in err.split("\n"):
        m = re.match(r"^\t(\d+):([\d,]+)-([\d,]+) \((.*)\) (\S+) (\(.\))$", line)
        if m:
            ref.append((int(m.group(1)), int(m.group(2).replace(",", "")), int(m.group(3).replace(",", "")), m.group(4), m.group(5), m.group(6)))
    ids = ann.ids()
    mine = [(int(ann.chr[i]), int(ann.start[i]), int(ann.end[i]), cfg.names[ann.type[i]], ids[i], "(+)" if ann.strand[i] == 1 else "(-)") for i in range(ann.n)]
    assert mine == ref, "host GTF front-end disagrees with the reference's interval list"
    hits, _ = host.read_hits(ann, BAM, "F")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "chrY.npz"),
                        f_chr=ann.chr, f_start=ann.start, f_end=ann.end, f_type=ann.type, f_strand=ann.strand, n_chr=np.uint32(ann.n_chr),
                        h_start=hits.start, h_end=hits.end, h_meta=hits.meta, h_nh=hits.nh, h_key=hits.read_key,
                        elem_line=cfg.elem_line, elem_strand=cfg.elem_strand, elem_vicinity=cfg.elem_vicinity)
    expected = {"config_text": open(CFG).read(), "element_names": cfg.names, "n_features": ann.n, "n_genes": ann.n_genes, "cases": {}}
    # a d/D case needs its own features: store them under a second key
    ann2 = host.Annotation(cfg, GTF, 5000, 200)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "chrY_d5000_D200.npz"),
                        f_chr=ann2.chr, f_start=ann2.start, f_end=ann2.end, f_type=ann2.type, f_strand=ann2.strand, n_chr=np.uint32(ann2.n_chr))
    with tempfile.TemporaryDirectory() as tmp:
        for name, kind, extra in MATRIX:
            mfile = os.path.join(tmp, name + ".m")
            args = ["-a", GTF, "-r", BAM, "-c", CFG] + [mfile if a == "@M" else a for a in extra]
            rc, out, err = pyoracle.run_reference(args, kind=kind)
            assert rc == 0, err
            _, rows = pyoracle.parse_table(out)
            case = {"args": extra, "binary": kind, "table": {k: v[0] for k, v in rows.items()}, "stats": pyoracle.parse_stats(err)[0]}
            if os.path.exists(mfile):
                case["read_stats_sorted"] = sorted(open(mfile).read().split("\n"))[:0]  # lines kept out of the repo; md5 below
                import hashlib
                case["read_stats_sorted_md5"] = hashlib.md5("\n".join(sorted(l for l in open(mfile).read().split("\n") if l)).encode()).hexdigest()
            expected["cases"][name] = case
    json.dump(expected, open(os.path.join(ROOT, "tests", "golden", "chrY_expected.json"), "w"), indent=1, sort_keys=True)
    print("golden fixtures written:", len(expected["cases"]), "cases,", hits.n, "hits,", ann.n, "features")


if __name__ == "__main__":
    main()
