import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from tests import fuzz
from oracle import pyoracle
from mmannot_b200 import device, host

def run(et, feats, hits, batch):
    a = device.Annotator(et, strategy="random", overlap=1.0, max_batch_hits=batch)
    a.load_features(feats); a.submit(0, hits); r = a.finish(0); a.close()
    return device.values_by_mask(r["rows"]), r["stats"]

found = 0
for seed in range(400):
    rng = np.random.default_rng(90000 + seed)
    et = fuzz.make_elements(rng, n_elements=3)
    feats = fuzz.make_features(rng, et, n_chr=1, n_feat=6, extent=300, max_len=200)
    n_reads = int(rng.integers(2, 12))
    hits = fuzz.make_hits(rng, feats, n_reads=n_reads, extent=300, max_nh=4, max_read=30, messy=0.3)
    ref = pyoracle.run(et.elem_line, et.elem_strand, et.elem_vicinity, feats, hits, strategy="random", overlap=1.0, want_hit_masks=True)
    got, st = run(et, feats, hits, 1 << 20)
    if got != ref["rows"]:
        found += 1
        print("seed", seed, "n", hits.n, "got", got, "want", ref["rows"])
        print("  keys", [int(k) % 1000 for k in hits.read_key], "nh", hits.nh.tolist(), "masks", [int(m) for m in ref["hit_mask"]])
        print("  rand", pyoracle.glibc_rand(1, 12).tolist())
        if found >= 3: break
print("searched; mismatches:", found)
