"""Read-key verification of the host decoder (XamReader::keyCollision): the device tells reads apart by a 64-bit key of the
name, the reference by the name string (mmannot.cpp:1656-1662, 1671).  Neighbouring records that share a key must share the
name; the first pair that does not is reported (and Counter::read refuses the file under -y default / random).  No pair of
real names collides in 64 bits, so the test knob MMANNOT_B200_KEY_MASK cuts the keys down to a few bits."""
import json
import os

import numpy as np
import pytest

from tests import common
from tests.test_sam_input import bam_records
from mmannot_b200 import host

CFGS = json.load(open(os.path.join(common.GOLDEN, "configs.json")))


@pytest.fixture(scope="module")
def sample(tmp_path_factory):
    common.ensure_built(("host",))
    d = tmp_path_factory.mktemp("keys")
    cfg = str(d / "c.txt")
    open(cfg, "w").write(CFGS["configTAIR10"])
    synth = host.Synth("tair10", 77, gene_scale=0.05, max_nh=6)
    gtf, bam = str(d / "a.gff"), str(d / "r.bam")
    synth.write_annotation(gtf)
    synth.write_bam(bam, 0, 30000)
    ann = host.Annotation(host.Config(cfg), gtf)
    return ann, bam, [r[0] for r in bam_records(bam)[1]]


def first_collision(names, mask):
    """Restatement: walk the records, remember the last name whose (masked) key differed from its predecessor's."""
    prev_name, prev_key = None, None
    for nm in names:
        k = host.name_key(nm) & mask
        if prev_name is None or k != prev_key:
            prev_name, prev_key = nm, k
        elif nm != prev_name:
            return "'%s' and '%s'" % (prev_name, nm)
    return ""


def decode(ann, bam, monkeypatch, mask=None, threads=None):
    if mask is None:
        monkeypatch.delenv("MMANNOT_B200_KEY_MASK", raising=False)
    else:
        monkeypatch.setenv("MMANNOT_B200_KEY_MASK", hex(mask))
    if threads is None:
        monkeypatch.delenv("MMANNOT_B200_DECODE_THREADS", raising=False)
    else:
        monkeypatch.setenv("MMANNOT_B200_DECODE_THREADS", str(threads))
    got = []
    hits, _ = host.read_hits(ann, bam, "F", collision=got)
    return hits, got[0]


def test_full_keys_never_collide(sample, monkeypatch):
    ann, bam, names = sample
    hits, col = decode(ann, bam, monkeypatch)
    assert hits.n == len(names)
    assert col == ""
    assert first_collision(names, (1 << 64) - 1) == ""


@pytest.mark.parametrize("mask", [0xFF, 0xFFF, 0xFFFF, 0x3])
@pytest.mark.parametrize("threads", [1, 6])
def test_masked_keys_report_the_first_neighbouring_pair(sample, monkeypatch, mask, threads):
    """threads = 6: the records of a chunk are parsed by clones side by side; the borders of their ranges are checked when the
    ranges are merged, so the answer does not depend on the number of threads."""
    ann, bam, names = sample
    hits, col = decode(ann, bam, monkeypatch, mask=mask, threads=threads)
    want = first_collision(names, mask)
    assert col == want
    if mask <= 0xFFF:
        assert want != ""  # 30 000 reads with keys of at most 12 bits: some neighbours share one
    keys = np.array([host.name_key(n) & mask for n in names], np.uint64)
    assert np.array_equal(hits.read_key, keys)
