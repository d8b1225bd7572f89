"""Host BAM decoder: the multi-threaded route (BGZF members inflated side by side, records parsed by parser clones on
contiguous ranges) must produce exactly the hits and warnings of the single-threaded one."""
import json
import os

import numpy as np
import pytest

from tests import common
from mmannot_b200 import host

CFGS = json.load(open(os.path.join(common.GOLDEN, "configs.json")))


@pytest.mark.parametrize("shape,cfg_key,spec", [("tair10", "configTAIR10", dict(max_nh=20)),
                                                ("flybase6", "configFlybase6", dict(max_nh=8, paired=True, rna_seq=True))])
def test_parallel_decode_equals_sequential(tmp_path, monkeypatch, shape, cfg_key, spec):
    cfg_path = str(tmp_path / "c.txt")
    open(cfg_path, "w").write(CFGS[cfg_key])
    synth = host.Synth(shape, 4321, gene_scale=0.05, **spec)
    gtf = str(tmp_path / "a.gtf")
    synth.write_annotation(gtf)
    # an annotation that does not know every chromosome of the reads: "unknown chromosome" warnings, once each, in order
    lines = [l for l in open(gtf) if not l.startswith(("Chr3", "3R", "chr3"))]
    open(gtf, "w").writelines(lines)
    ann = host.Annotation(host.Config(cfg_path), gtf)
    bam = str(tmp_path / "r.bam")
    synth.write_bam(bam, 0, 40000)
    out = {}
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("MMANNOT_B200_DECODE_THREADS", threads)
        out[threads] = host.read_hits(ann, bam, "F")
    for threads in ("3", "8"):
        for k in ("start", "end", "meta", "nh", "read_key"):
            assert np.array_equal(getattr(out["1"][0], k), getattr(out[threads][0], k)), (threads, k)
        assert out["1"][1] == out[threads][1]
    assert out["1"][0].n > 40000
    assert "is not present in your annotation file" in out["1"][1]
