"""Pins the CPU oracle (and the host front-end that produces its buffers) against the REFERENCE ITSELF on name-grouped
multi-mapping input: synthetic annotation + BAM of the benchmark shapes, run through oracle/_ref/mmannot_fixed
(the reference compiled where it lies, see oracle/build_ref.sh) and through oracle.c on the hits our decoder
extracts from the same BAM.  Skipped where the compiled reference is not present."""
import json
import os

import numpy as np
import pytest

from tests import common
from oracle import pyoracle
from mmannot_b200 import host
from mmannot_b200.device import round_half_away

pytestmark = pytest.mark.skipif(pyoracle.ref_binary("fixed") is None, reason="oracle/_ref not built (needs /root/reference)")

CFGS = json.load(open(os.path.join(common.GOLDEN, "configs.json")))

SHAPES = [("tair10", "configTAIR10", dict(max_nh=20)), ("hs38", "configHS38", dict(max_nh=30)),
          ("flybase6", "configFlybase6", dict(max_nh=8, paired=True, rna_seq=True))]


@pytest.fixture(scope="module", autouse=True)
def _build():
    common.ensure_built()


@pytest.fixture(scope="module", params=SHAPES, ids=[s[0] for s in SHAPES])
def workload(request, tmp_path_factory):
    shape, cfg_key, spec = request.param
    tmp = tmp_path_factory.mktemp(shape)
    cfg_path = str(tmp / (cfg_key + ".txt"))
    open(cfg_path, "w").write(CFGS[cfg_key])
    synth = host.Synth(shape, 4242, gene_scale=0.03, **spec)
    gtf = str(tmp / "a.gtf")
    synth.write_annotation(gtf)
    bam = str(tmp / "reads.bam")
    synth.write_bam(bam, 0, 6000)
    cfg = host.Config(cfg_path)
    ann = host.Annotation(cfg, gtf)
    return dict(cfg_path=cfg_path, gtf=gtf, bam=bam, cfg=cfg, ann=ann, synth=synth)


CASES = [(["-s", "F"], {}), (["-s", "R", "-l", "1"], {}), (["-s", "U", "-l", "0.5"], {}), (["-s", "F", "-l", "15"], {}),
         (["-s", "F", "-y", "unique"], {}), (["-s", "F", "-y", "ratio"], {}), (["-s", "F", "-y", "random"], {}),
         (["-s", "U", "-d", "300", "-D", "2500"], dict(up=300, down=2500))]


@pytest.mark.parametrize("args,ann_opt", CASES, ids=[" ".join(c[0]) for c in CASES])
def test_oracle_equals_reference(workload, args, ann_opt):
    w = workload
    rc, out, err = pyoracle.run_reference(["-a", w["gtf"], "-r", w["bam"], "-c", w["cfg_path"]] + args, kind="fixed")
    assert rc == 0, err
    _, ref_rows = pyoracle.parse_table(out)
    ref_stats = pyoracle.parse_stats(err)[0]
    o = common.case_options(args)
    ann = w["ann"] if not ann_opt else host.Annotation(w["cfg"], w["gtf"], ann_opt["up"], ann_opt["down"])
    hits, _ = host.read_hits(ann, w["bam"], o["strand"])
    # the packed hits decoded from the BAM are the ones the generator produces directly (bench.py relies on that)
    direct = w["synth"].hits(ann, o["strand"], 0, 6000)
    for k in ("start", "end", "meta", "nh", "read_key"):
        assert np.array_equal(getattr(hits, k), getattr(direct, k)), k
    res = pyoracle.run(w["cfg"].elem_line, w["cfg"].elem_strand, w["cfg"].elem_vicinity, ann, hits, strategy=o["strategy"], overlap=o["overlap"])
    table = {w["cfg"].row_name(m): round_half_away(v) for m, v in res["rows"].items()}
    assert table == {k: v[0] for k, v in ref_rows.items()}
    for k, v in ref_stats.items():
        assert res["stats"][k] == v, k


@pytest.mark.parametrize("args", [["-s", "F", "-y", "ratio"], ["-s", "U", "-y", "ratio", "-l", "0.5"]], ids=["F", "U -l 0.5"])
def test_ratio_table_from_integer_counts_equals_reference(workload, args):
    """-y ratio: the device keeps INTEGER counts per (element set, NH) and the host forms each cell as the sum over NH, in
    ascending NH, of count * (1.0 / NH) (Counter::read, csrc/host/counter.cpp), where the reference adds 1.0 / NH hit by hit in
    file order (mm:1730) and prints (unsigned) round(value) (mm:1868).  The two double sums may differ in the last bits; the
    printed tables must not.  Here the integer counts come from the oracle's per-hit element sets."""
    w = workload
    rc, out, err = pyoracle.run_reference(["-a", w["gtf"], "-r", w["bam"], "-c", w["cfg_path"]] + args, kind="fixed")
    assert rc == 0, err
    _, ref_rows = pyoracle.parse_table(out)
    o = common.case_options(args)
    hits, _ = host.read_hits(w["ann"], w["bam"], o["strand"])
    res = pyoracle.run(w["cfg"].elem_line, w["cfg"].elem_strand, w["cfg"].elem_vicinity, w["ann"], hits, strategy="ratio",
                       overlap=o["overlap"], want_hit_masks=True)
    masks, nh = res["hit_mask"], hits.nh
    counts = {}
    for m, n in zip(masks[masks != 0].tolist(), nh[masks != 0].tolist()):
        counts[(m, n)] = counts.get((m, n), 0) + 1
    cells = {}
    for (m, n) in sorted(counts):  # ascending (set, NH), like the sort of the rows in Counter::read
        cells[m] = cells.get(m, 0.0) + float(counts[(m, n)]) * (1.0 / n if n else 1.0)
    table = {w["cfg"].row_name(m): round_half_away(v) for m, v in cells.items()}
    assert table == {k: v[0] for k, v in ref_rows.items()}
    # and the doubles themselves agree with the hit-by-hit sums within the stated tolerance
    for m, v in res["rows"].items():
        assert abs(cells[m] - v) <= 1e-9 * max(1.0, abs(v))
