"""Shared fixtures / helpers of the test-suite."""
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

from mmannot_b200 import host  # noqa: E402


def ensure_built(targets=("host", "oracle/_build/liboracle.so")):
    subprocess.check_call(["make", "-s"] + list(targets), cwd=ROOT)


def load_chrY(variant="chrY"):
    z = np.load(os.path.join(GOLDEN, "chrY.npz"))
    exp = json.load(open(os.path.join(GOLDEN, "chrY_expected.json")))
    et = host.ElementTable(z["elem_line"], z["elem_strand"], z["elem_vicinity"], exp["element_names"])
    fz = z if variant == "chrY" else np.load(os.path.join(GOLDEN, variant + ".npz"))
    feats = host.FeatureArrays(fz["f_chr"], fz["f_start"], fz["f_end"], fz["f_type"], fz["f_strand"], int(fz["n_chr"]))
    hits = host.Hits(z["h_start"], z["h_end"], z["h_meta"], z["h_nh"], z["h_key"])  # decoded with -s F
    return et, feats, hits, exp


def restrand(hits, s):
    """Hits decoded with -s F -> the same hits under -s U / -s R (strand mapping mm:836-844)."""
    meta = hits.meta.copy()
    if s == "U":
        meta |= np.uint32(0x80000000)
    elif s == "R":
        meta ^= np.uint32(0x80000000)
    return host.Hits(hits.start, hits.end, meta, hits.nh, hits.read_key)


def case_options(args):
    """Reference command-line arguments of a golden case -> keyword options of the hot path."""
    o = {"strand": "F", "strategy": "default", "overlap": -1.0, "rescue_threshold": 1.0, "read_stats": False, "variant": "chrY"}
    i = 0
    while i < len(args):
        a = args[i]
        if a == "-s": o["strand"] = args[i + 1]
        elif a == "-y": o["strategy"] = args[i + 1]
        elif a == "-l": o["overlap"] = float(np.float32(float(args[i + 1])))
        elif a == "-e": o["rescue_threshold"] = float(np.float32(float(np.float32(float(args[i + 1]))) / 100.0))  # (float)(stof(arg) / 100.0), mm:2024
        elif a == "-m": o["read_stats"] = True
        elif a in ("-d", "-D"): o["variant"] = "chrY_d5000_D200"
        i += 2
    return o


def table_of(et, values_by_mask):
    from mmannot_b200.device import round_half_away
    return {et.row_name(m): round_half_away(v) for m, v in values_by_mask.items()}
