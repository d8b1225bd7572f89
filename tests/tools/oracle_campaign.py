"""One-off campaign (not part of the suite): the CPU oracle against the compiled reference (oracle/_ref/mmannot_fixed) over
many seeds x shapes x option sets x record orders, larger than tests/test_oracle_vs_reference_synth.py.
Usage: python tests/tools/oracle_campaign.py [n_seeds] [reads] [only]    (needs /root/reference to have been compiled by `make oracle`)"""
import itertools
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import common  # noqa: E402
from oracle import pyoracle  # noqa: E402
from mmannot_b200 import host  # noqa: E402
from mmannot_b200.device import round_half_away  # noqa: E402

CFGS = json.load(open(os.path.join(common.GOLDEN, "configs.json")))
SHAPES = [("tair10", "configTAIR10", dict(max_nh=20)), ("hs38", "configHS38", dict(max_nh=60)),
          ("flybase6", "configFlybase6", dict(max_nh=8, paired=True, rna_seq=True))]
OPTS = [["-s", "F"], ["-s", "R", "-l", "1"], ["-s", "U", "-l", "0.5"], ["-s", "F", "-l", "15"], ["-s", "R", "-l", "0.9"],
        ["-s", "F", "-y", "unique"], ["-s", "U", "-y", "ratio"], ["-s", "F", "-y", "random"], ["-s", "U", "-d", "300", "-D", "2500"],
        ["-s", "F", "-l", "0.99", "-y", "ratio"], ["-s", "F", "-m", "@", "-e", "50"], ["-s", "U", "-m", "@", "-e", "33"], ["-s", "R", "-m", "@", "-e", "80", "-y", "ratio"],
        ["-s", "F", "-m", "@", "-e", "67", "-y", "random"]]  # ("@" = a scratch file: the rescue test only runs with -m, mm:491)


def main():
    n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    reads = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    only = sys.argv[3] if len(sys.argv) > 3 else ""  # e.g. "ratio": only the option sets that contain the word
    bad = total = 0
    for seed, (shape, cfg_key, spec), cs in itertools.product(range(n_seeds), SHAPES, (False, True)):
        tmp = tempfile.mkdtemp(prefix="camp_")
        cfg_path = os.path.join(tmp, cfg_key + ".txt")
        open(cfg_path, "w").write(CFGS[cfg_key])
        synth = host.Synth(shape, 9000 + 17 * seed, gene_scale=(0.02, 0.05, 0.1, 0.01)[seed % 4], **spec)
        gtf, bam = os.path.join(tmp, "a.gtf"), os.path.join(tmp, "r.bam")
        synth.write_annotation(gtf)
        synth.write_bam(bam, 1000 * seed, reads, coordinate_sorted=cs)
        cfg = host.Config(cfg_path)
        for args in OPTS:
            if only and only not in args:
                continue
            if cs and "random" in args:
                continue  # (order-dependent draws: same stream, but the campaign keeps to what the suite pins)
            o = common.case_options(args)
            up, down = (300, 2500) if "-d" in args else (None, None)
            ann = host.Annotation(cfg, gtf, up, down) if up else host.Annotation(cfg, gtf)
            rc, out, err = pyoracle.run_reference(["-a", gtf, "-r", bam, "-c", cfg_path] + [os.path.join(tmp, "m.txt") if a == "@" else a for a in args], kind="fixed")
            assert rc == 0, err
            _, ref_rows = pyoracle.parse_table(out)
            ref_stats = pyoracle.parse_stats(err)[0]
            hits, _ = host.read_hits(ann, bam, o["strand"])
            res = pyoracle.run(cfg.elem_line, cfg.elem_strand, cfg.elem_vicinity, ann, hits, strategy=o["strategy"], overlap=o["overlap"],
                               rescue_threshold=o["rescue_threshold"], read_stats=o["read_stats"], want_hit_masks=(o["strategy"] == "ratio"))
            table = {cfg.row_name(m): round_half_away(v) for m, v in res["rows"].items()}
            ok = table == {k: v[0] for k, v in ref_rows.items()} and all(res["stats"][k] == v for k, v in ref_stats.items())
            if o["strategy"] == "ratio" and not o["read_stats"]:  # the cells as Counter::read forms them from the device's integer counts per (set, NH)
                masks, counts, cells = res["hit_mask"], {}, {}
                for m, n in zip(masks[masks != 0].tolist(), hits.nh[masks != 0].tolist()):
                    counts[(m, n)] = counts.get((m, n), 0) + 1
                for (m, n) in sorted(counts):
                    cells[m] = cells.get(m, 0.0) + float(counts[(m, n)]) * (1.0 / n if n else 1.0)
                ok = ok and {cfg.row_name(m): round_half_away(v) for m, v in cells.items()} == {k: v[0] for k, v in ref_rows.items()}
            total += 1
            if not ok:
                bad += 1
                print("DIFF", shape, "seed", seed, "coordinate-sorted" if cs else "name-grouped", " ".join(args), flush=True)
        print("done", shape, seed, "cs" if cs else "grouped", "cases so far", total, "bad", bad, flush=True)
    print("campaign: %d cases, %d differ" % (total, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
