"""GPU path on the benchmark shapes: (1) device == oracle on synthetic sRNA-Seq / RNA-Seq hits at a size the oracle
finishes in seconds, for several batch sizes; (2) the drop-in command line (mmannot_b200/bin/mmannot_b200: BAM decode
on the host, annotation on the GPU through the C ABI) writes the same table and counters as the reference binary
(oracle/_ref/mmannot_fixed, compiled from the reference where it lies) on the same files."""
import json
import os
import subprocess

import numpy as np
import pytest

from tests import common
from oracle import pyoracle
from mmannot_b200 import host

pytestmark = pytest.mark.gpu

CFGS = json.load(open(os.path.join(common.GOLDEN, "configs.json")))
SHAPES = [("tair10", "configTAIR10", dict(max_nh=20)), ("hs38", "configHS38", dict(max_nh=100)),
          ("flybase6", "configFlybase6", dict(max_nh=8, paired=True, rna_seq=True))]
CLI = os.path.join(common.ROOT, "mmannot_b200", "bin", "mmannot_b200")


@pytest.fixture(scope="module", params=SHAPES, ids=[s[0] for s in SHAPES])
def workload(request, tmp_path_factory):
    shape, cfg_key, spec = request.param
    tmp = tmp_path_factory.mktemp(shape)
    cfg_path = str(tmp / (cfg_key + ".txt"))
    open(cfg_path, "w").write(CFGS[cfg_key])
    synth = host.Synth(shape, 777, gene_scale=0.1, **spec)
    gtf = str(tmp / "a.gtf")
    synth.write_annotation(gtf)
    cfg = host.Config(cfg_path)
    ann = host.Annotation(cfg, gtf)
    return dict(tmp=tmp, cfg_path=cfg_path, gtf=gtf, cfg=cfg, ann=ann, synth=synth)


def _device(cfg, ann, hits, batch, **kw):
    from mmannot_b200 import device
    a = device.Annotator(cfg, max_batch_hits=batch, **kw)
    try:
        a.load_features(ann)
        a.submit(0, hits)
        r = a.finish(0)
        r["values"] = device.values_by_mask(r["rows"])
        return r
    finally:
        a.close()


@pytest.mark.parametrize("strategy,overlap", [("default", -1.0), ("default", 10.0), ("ratio", -1.0), ("unique", 0.5), ("random", -1.0)])
def test_device_equals_oracle_on_benchmark_shapes(workload, strategy, overlap):
    w = workload
    hits = w["synth"].hits(w["ann"], "F", 0, 150000, threads=4)
    ref = pyoracle.run(w["cfg"].elem_line, w["cfg"].elem_strand, w["cfg"].elem_vicinity, w["ann"], hits, strategy=strategy, overlap=overlap)
    for batch in (50021, 1 << 22):
        res = _device(w["cfg"], w["ann"], hits, batch, strategy=strategy, overlap=overlap)
        assert set(res["values"]) == set(ref["rows"])
        for m, v in ref["rows"].items():
            if strategy == "ratio":
                assert abs(res["values"][m] - v) <= 1e-9 * max(1.0, abs(v))
            else:
                assert res["values"][m] == v
        assert res["stats"] == ref["stats"]


def test_coordinate_sorted_variant(workload):
    """Same reads in coordinate order: the records of a read are scattered, everything takes the deferred path."""
    w = workload
    hits = w["synth"].hits(w["ann"], "F", 0, 60000, threads=4)
    order = np.lexsort((hits.start, hits.meta & np.uint32(0xFFFFFF)))
    shuffled = host.Hits(hits.start[order], hits.end[order], hits.meta[order], hits.nh[order], hits.read_key[order])
    ref = pyoracle.run(w["cfg"].elem_line, w["cfg"].elem_strand, w["cfg"].elem_vicinity, w["ann"], shuffled)
    res = _device(w["cfg"], w["ann"], shuffled, 30011)
    assert res["values"] == ref["rows"] and res["stats"] == ref["stats"]


@pytest.mark.skipif(pyoracle.ref_binary("fixed") is None or not os.path.exists(CLI), reason="needs oracle/_ref and the CLI binary")
@pytest.mark.parametrize("args", [["-s", "F"], ["-s", "U", "-l", "1"], ["-s", "R", "-y", "ratio"], ["-s", "F", "-y", "unique", "-l", "0.5"],
                                  ["-s", "F", "-y", "random"]],  # two files: the rand() stream runs on from the first into the second
                         ids=lambda a: " ".join(a))
def test_cli_equals_reference(workload, args):
    w = workload
    bams = []
    for i in range(2):
        b = str(w["tmp"] / ("sample%d.bam" % i))
        if not os.path.exists(b):
            w["synth"].write_bam(b, i * 20000, 20000)
        bams.append(b)
    common_args = ["-a", w["gtf"], "-c", w["cfg_path"], "-r"] + bams + args
    rc, ref_out, ref_err = pyoracle.run_reference(common_args, kind="fixed")
    assert rc == 0, ref_err
    pr = subprocess.run([CLI] + common_args, capture_output=True, text=True, timeout=300)
    assert pr.returncode == 0, pr.stderr
    assert pr.stdout == ref_out
    assert pyoracle.parse_stats(pr.stderr) == pyoracle.parse_stats(ref_err)
    # same report text for the per-sample blocks
    pick = lambda t: [l for l in t.split("\n") if l.startswith("\t#") or l.startswith("Results for")]
    assert pick(pr.stderr) == pick(ref_err)


STATS_CASES = [["-s", "F"], ["-s", "F", "-e", "60"], ["-s", "U", "-l", "1", "-e", "75"], ["-s", "R", "-y", "ratio", "-e", "50"],
               ["-s", "F", "-y", "unique"], ["-s", "F", "-y", "random", "-e", "80"]]


@pytest.mark.skipif(pyoracle.ref_binary("fixed") is None or not os.path.exists(CLI), reason="needs oracle/_ref and the CLI binary")
@pytest.mark.parametrize("args", STATS_CASES, ids=lambda a: " ".join(a))
@pytest.mark.parametrize("sorted_input", [False, True], ids=["grouped", "coordinate-sorted"])
def test_cli_read_and_interval_statistics(workload, args, sorted_input):
    """-m (per read) and -M (per interval) files, the table and the counters against the reference binary, byte for byte.
    With -m the -e threshold becomes active (mmannot.cpp:491).  The coordinate-sorted input leaves many reads open until
    the end of the file: their -m lines come out in the iteration order of the reference's name-keyed map."""
    w = workload
    tag = "s" if sorted_input else "g"
    bam = str(w["tmp"] / ("stats_%s.bam" % tag))
    if not os.path.exists(bam):
        w["synth"].write_bam(bam, 50000, 6000, coordinate_sorted=sorted_input)
    out = {}
    for who, exe in (("ref", None), ("b200", CLI)):
        m, M = str(w["tmp"] / ("%s_m.txt" % who)), str(w["tmp"] / ("%s_M.txt" % who))
        a = ["-a", w["gtf"], "-c", w["cfg_path"], "-r", bam, "-m", m, "-M", M] + args
        if exe is None:
            rc, so, se = pyoracle.run_reference(a, kind="fixed")
        else:
            pr = subprocess.run([exe] + a, capture_output=True, text=True, timeout=300)
            rc, so, se = pr.returncode, pr.stdout, pr.stderr
        assert rc == 0, se
        out[who] = (so, pyoracle.parse_stats(se), open(m).read(), open(M).read())
    assert out["b200"][0] == out["ref"][0]
    assert out["b200"][1] == out["ref"][1]
    assert out["b200"][3] == out["ref"][3], "-M differs"
    assert out["b200"][2] == out["ref"][2], "-m differs"
    assert len(out["ref"][2]) > 1000 and len(out["ref"][3]) > 1000


@pytest.mark.skipif(pyoracle.ref_binary("fixed") is None or not os.path.exists(CLI), reason="needs oracle/_ref and the CLI binary")
def test_cli_threads_spread_files_over_contexts(workload):
    """-t n: n workers, one context each (on as many GPUs as there are), files in turn; same table and per-file reports,
    in file order, as the reference run one file after the other."""
    w = workload
    bams = []
    for i in range(3):
        b = str(w["tmp"] / ("t%d.bam" % i))
        if not os.path.exists(b):
            w["synth"].write_bam(b, 100000 + i * 7000, 7000)
        bams.append(b)
    common_args = ["-a", w["gtf"], "-c", w["cfg_path"], "-r"] + bams + ["-s", "F", "-y", "ratio"]
    rc, ref_out, ref_err = pyoracle.run_reference(common_args, kind="fixed")
    assert rc == 0, ref_err
    for t in ("2", "3", "8"):
        pr = subprocess.run([CLI] + common_args + ["-t", t], capture_output=True, text=True, timeout=300)
        assert pr.returncode == 0, pr.stderr
        assert pr.stdout == ref_out
        pick = lambda txt: [l for l in txt.split("\n") if l.startswith("\t#") or l.startswith("Results for")]
        assert pick(pr.stderr) == pick(ref_err)


@pytest.mark.skipif(pyoracle.ref_binary("fixed") is None or not os.path.exists(CLI), reason="needs oracle/_ref and the CLI binary")
def test_random_remembers_read_names_across_input_files(workload):
    """-y random keeps seen / chosenId / numberSeen over the input files of a run (Counter::clear does not touch them,
    mmannot.cpp:1742-1747): a name that was counted in one file is ignored in the next, a name whose drawn hit lies beyond its
    hits in one file goes on counting in the next.  Three files with overlapping read ranges (so that names repeat), also
    decoded on the host."""
    w = workload
    bams = []
    for i, (first, n) in enumerate(((0, 12000), (6000, 12000), (0, 20000))):
        b = str(w["tmp"] / ("overlap%d.bam" % i))
        if not os.path.exists(b):
            w["synth"].write_bam(b, first, n)
        bams.append(b)
    common_args = ["-a", w["gtf"], "-c", w["cfg_path"], "-r"] + bams + ["-s", "F", "-y", "random"]
    rc, ref_out, ref_err = pyoracle.run_reference(common_args, kind="fixed")
    assert rc == 0, ref_err
    for env in ({}, {"MMANNOT_B200_HOST_DECODE": "1"}):
        pr = subprocess.run([CLI] + common_args, capture_output=True, text=True, timeout=300, env=dict(os.environ, **env))
        assert pr.returncode == 0, pr.stderr
        assert pr.stdout == ref_out
        assert pyoracle.parse_stats(pr.stderr) == pyoracle.parse_stats(ref_err)
