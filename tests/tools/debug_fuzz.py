import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from tests import fuzz
from oracle import pyoracle
from mmannot_b200 import device

for seed in range(6):
  for overlap in (-1.0, 0.5, 1.0, 12.0):
    rng = np.random.default_rng(1000 + seed)
    et = fuzz.make_elements(rng)
    feats = fuzz.make_features(rng, et, n_feat=int(rng.integers(50, 900)))
    hits = fuzz.make_hits(rng, feats, n_reads=20000, max_nh=1, messy=0.0)
    ref = pyoracle.run(et.elem_line, et.elem_strand, et.elem_vicinity, feats, hits, overlap=overlap, want_hit_masks=True)
    for fast in (None, 0, 3, 9):
        a = device.Annotator(et, overlap=overlap, fast_bin_shift=fast, max_batch_hits=1 << 20)
        a.load_features(feats)
        got = a.annotate(hits)
        a.close()
        bad = np.nonzero(got != ref["hit_mask"])[0]
        if len(bad):
            print("seed", seed, "overlap", overlap, "fast", fast, "mismatches", len(bad))
            print(" elem line", et.elem_line, "strand", et.elem_strand, "vic", et.elem_vicinity)
            for i in bad[:3]:
                c = int(hits.meta[i] & 0xFFFFFF); rs = int(hits.start[i]); re = int(hits.end[i])
                print("  hit", i, "chr", c, "rs", rs, "re", re, "strand", int(hits.meta[i] >> 31), "got", hex(int(got[i])), "want", hex(int(ref["hit_mask"][i])))
                sel = np.nonzero((feats.chr == c) & (feats.start <= re + 2) & (feats.end >= rs - 2))[0]
                for v in sel:
                    print("     feat", v, int(feats.start[v]), int(feats.end[v]), "type", int(feats.type[v]), "strand", int(feats.strand[v]))
print("done")
