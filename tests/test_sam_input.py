"""SAM text input (SamReader, mmannot.cpp:1431-1479) with XA alternate hits (mmannot.cpp:1360-1399): the reference binary on
a SAM file against the oracle on the hits OUR decoder extracts from the same file.  The SAM is made from a synthetic BAM by a
small converter below; a third of its records get an XA:Z tag listing 1-3 alternate placements."""
import gzip
import json
import os
import struct

import numpy as np
import pytest

from tests import common
from oracle import pyoracle
from mmannot_b200 import host
from mmannot_b200.device import round_half_away

pytestmark = pytest.mark.skipif(pyoracle.ref_binary("fixed") is None, reason="oracle/_ref not built (needs /root/reference)")
CFGS = json.load(open(os.path.join(common.GOLDEN, "configs.json")))


def bam_records(path):
    """(name, flag, rname, pos1, cigar text, NH) of every record of a BAM file (BGZF = concatenated gzip members)."""
    data = gzip.open(path, "rb").read()
    assert data[:4] == b"BAM\1"
    l_text, = struct.unpack_from("<i", data, 4)
    p = 8 + l_text
    n_ref, = struct.unpack_from("<i", data, p); p += 4
    refs = []
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", data, p); p += 4
        refs.append(data[p:p + l_name - 1].decode()); p += l_name + 4
    out = []
    while p < len(data):
        bs, = struct.unpack_from("<i", data, p); p += 4
        ref_id, pos, l_rn, mapq, _bin, n_cig, flag, l_seq = struct.unpack_from("<iiBBHHHi", data, p)
        q = p + 32
        name = data[q:q + l_rn - 1].decode(); q += l_rn
        cigar = ""
        for _ in range(n_cig):
            v, = struct.unpack_from("<I", data, q); q += 4
            cigar += "%d%s" % (v >> 4, "MIDNSHP=X"[v & 15])
        q += (l_seq + 1) // 2 + l_seq
        nh = 1
        end = p + bs
        while q + 3 <= end:
            tag, ty = data[q:q + 2], chr(data[q + 2]); q += 3
            size = {"A": 1, "c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}[ty]
            val = int.from_bytes(data[q:q + size], "little"); q += size
            if tag == b"NH":
                nh = val
        out.append((name, flag, refs[ref_id] if ref_id >= 0 else "*", pos + 1, cigar or "*", nh))
        p = end
    return refs, out


def write_sam(path, refs, records, rng, with_xa):
    with open(path, "w") as f:
        f.write("@HD\tVN:1.0\tSO:unsorted\n")
        for r in refs:
            f.write("@SQ\tSN:%s\tLN:100000000\n" % r)
        for i, (name, flag, rname, pos, cigar, nh) in enumerate(records):
            tags = ["NM:i:0"]
            if with_xa and rng.random() < 0.33:
                alts = []
                for _ in range(int(rng.integers(1, 4))):
                    o = records[int(rng.integers(0, len(records)))]
                    alts.append("%s,%s%d,%s,0" % (o[2], "-" if o[1] & 16 else "+", o[3], o[4]))
                tags.append("XA:Z:" + ";".join(alts) + ";")
            else:
                tags.append("NH:i:%d" % nh)
            f.write("\t".join([name, str(flag), rname, str(pos), "255", cigar, "*", "0", "0", "*", "*"] + tags) + "\n")


@pytest.mark.parametrize("with_xa", [False, True], ids=["plain", "XA"])
@pytest.mark.parametrize("args", [["-s", "F"], ["-s", "U", "-l", "1", "-y", "ratio"], ["-s", "R", "-y", "unique"]], ids=lambda a: " ".join(a))
def test_sam_text_input(tmp_path, args, with_xa):
    common.ensure_built()
    cfg_path = str(tmp_path / "c.txt")
    open(cfg_path, "w").write(CFGS["configTAIR10"])
    synth = host.Synth("tair10", 2468, gene_scale=0.03, max_nh=6)
    gtf = str(tmp_path / "a.gtf"); synth.write_annotation(gtf)
    bam = str(tmp_path / "r.bam"); synth.write_bam(bam, 0, 3000)
    refs, records = bam_records(bam)
    sam = str(tmp_path / "r.sam")
    write_sam(sam, refs, records, np.random.default_rng(7), with_xa)
    cfg = host.Config(cfg_path); ann = host.Annotation(cfg, gtf)
    rc, out, err = pyoracle.run_reference(["-a", gtf, "-r", sam, "-c", cfg_path] + args, kind="fixed")
    assert rc == 0, err
    _, ref_rows = pyoracle.parse_table(out)
    ref_stats = pyoracle.parse_stats(err)[0]
    o = common.case_options(args)
    hits, warn = host.read_hits(ann, sam, o["strand"])
    if not with_xa:  # the SAM carries exactly the records of the BAM
        hb, _ = host.read_hits(ann, bam, o["strand"])
        for k in ("start", "end", "meta", "nh", "read_key"):
            assert np.array_equal(getattr(hits, k), getattr(hb, k)), k
    else:
        assert hits.n > len(records)  # the alternates are hits of their own
    res = pyoracle.run(cfg.elem_line, cfg.elem_strand, cfg.elem_vicinity, ann, hits, strategy=o["strategy"], overlap=o["overlap"])
    table = {cfg.row_name(m): round_half_away(v) for m, v in res["rows"].items()}
    assert table == {k: v[0] for k, v in ref_rows.items()}
    for k, v in ref_stats.items():
        assert res["stats"][k] == v, k
