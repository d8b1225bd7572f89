"""GPU debugging aid: device vs oracle on a synthetic benchmark shape; per-hit masks first, then the smallest failing prefix."""
import sys, os, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from tests import common
from oracle import pyoracle
from mmannot_b200 import device, host

shape = sys.argv[1] if len(sys.argv) > 1 else "hs38"
cfgkey = {"tair10": "configTAIR10", "hs38": "configHS38", "flybase6": "configFlybase6"}[shape]
spec = {"tair10": dict(max_nh=20), "hs38": dict(max_nh=100), "flybase6": dict(max_nh=8, paired=True, rna_seq=True)}[shape]
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 150000
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 22
CFGS = json.load(open(os.path.join(common.GOLDEN, "configs.json")))
tmp = tempfile.mkdtemp()
cfg_path = os.path.join(tmp, cfgkey + ".txt"); open(cfg_path, "w").write(CFGS[cfgkey])
synth = host.Synth(shape, 777, gene_scale=0.1, **spec)
gtf = os.path.join(tmp, "a.gtf"); synth.write_annotation(gtf)
cfg = host.Config(cfg_path); ann = host.Annotation(cfg, gtf)
hits = synth.hits(ann, "F", 0, n_reads, threads=4)
print("hits", hits.n, "features", ann.n, "E", cfg.n_elements)
ref = pyoracle.run(cfg.elem_line, cfg.elem_strand, cfg.elem_vicinity, ann, hits, want_hit_masks=True)
a = device.Annotator(cfg, max_batch_hits=1 << 22); a.load_features(ann)
got = a.annotate(hits); a.close()
bad = np.nonzero(got != ref["hit_mask"])[0]
print("per-hit mask mismatches:", len(bad))
for i in bad[:5]:
    print("  hit", i, "chr", int(hits.meta[i] & 0xFFFFFF), int(hits.start[i]), int(hits.end[i]), "got", hex(int(got[i])), "want", hex(int(ref["hit_mask"][i])))

def run_dev(h, batch):
    a = device.Annotator(cfg, max_batch_hits=batch)
    try:
        a.load_features(ann); a.submit(0, h); r = a.finish(0)
    finally:
        a.close()
    return device.values_by_mask(r["rows"]), r["stats"]

def ok(n):
    h = hits.slice(0, n)
    r = pyoracle.run(cfg.elem_line, cfg.elem_strand, cfg.elem_vicinity, ann, h)
    v, s = run_dev(h, batch)
    return v == r["rows"] and s == r["stats"], v, s, r

# read boundaries (name-grouped)
heads = np.nonzero(np.concatenate([[True], hits.read_key[1:] != hits.read_key[:-1]]))[0]
good, v, s, r = ok(hits.n)
print("full:", "OK" if good else "MISMATCH", s, r["stats"])
if not good:
    diff = {m: (v.get(m), r["rows"].get(m)) for m in set(v) | set(r["rows"]) if v.get(m) != r["rows"].get(m)}
    print("row diffs (device, oracle):", {hex(k): x for k, x in list(diff.items())[:10]})
    lo, hi = 0, len(heads)  # number of reads in the prefix; lo ok, hi bad
    while hi - lo > 1:
        mid = (lo + hi) // 2
        n = int(heads[mid]) if mid < len(heads) else hits.n
        if ok(n)[0]: lo = mid
        else: hi = mid
    n_lo = int(heads[lo]); n_hi = int(heads[hi]) if hi < len(heads) else hits.n
    print("smallest failing prefix: reads", hi, "hits", n_hi, "; last read = records", n_lo, "..", n_hi - 1)
    for i in range(max(0, n_lo - 3), min(hits.n, n_hi + 2)):
        print("   rec", i, "key", hex(int(hits.read_key[i]))[-8:], "nh", int(hits.nh[i]), "chr", int(hits.meta[i] & 0xFFFFFF), int(hits.start[i]), int(hits.end[i]),
              "mask", hex(int(ref["hit_mask"][i])), "tilepos", i % 128, "tile", i // 128)
    g, v, s, r = ok(n_hi)
    diff = {m: (v.get(m), r["rows"].get(m)) for m in set(v) | set(r["rows"]) if v.get(m) != r["rows"].get(m)}
    print("prefix row diffs (device, oracle):", {hex(k): x for k, x in diff.items()}, s, r["stats"])
