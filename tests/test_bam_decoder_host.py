"""The device BAM decoder's inflate and record functions (mma_bam.cuh; __host__ __device__) run on the CPU: every BGZF member
inflated by inflateMember must equal zlib's output -- fixed, dynamic and stored deflate blocks, members recompressed at other
levels -- and the hits parsed from the records must equal the host decoder's (XamReader) for every -s."""
import gzip
import json
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

from tests import common
from mmannot_b200 import host

CFGS = json.load(open(os.path.join(common.GOLDEN, "configs.json")))
TOOL_SRC = os.path.join(common.ROOT, "tests", "tools", "bam_host_check.cu")


def build_tool(exe, defines=()):
    nvcc = "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    # (the device side of the headers is compiled too, for the product's architecture; only the host side runs here)
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I" + os.path.join(common.ROOT, "include"),
                           "-I" + os.path.join(common.ROOT, "mmannot_b200", "csrc")] + list(defines) + ["-o", exe, TOOL_SRC, "-lz"])
    return exe


@pytest.fixture(scope="module")
def tool(tmp_path_factory):
    common.ensure_built(("host",))
    return build_tool(str(tmp_path_factory.mktemp("tool") / "bam_host_check"))


def recompress(src, dst, level):
    """Same members, deflated again at `level` (0 = stored blocks, 9 = long dynamic codes, Z_FIXED-like small inputs)."""
    data = open(src, "rb").read()
    out = bytearray()
    at = 0
    while at + 18 <= len(data):
        xlen = struct.unpack_from("<H", data, at + 10)[0]
        total = struct.unpack_from("<H", data, at + 16)[0] + 1
        raw = zlib.decompress(data[at + 12 + xlen:at + total - 8], -15)
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        body = co.compress(raw) + co.flush()
        if len(body) + 26 > 65536:  # (stored 64 KB members would not fit the 16-bit size field: keep the original)
            out += data[at:at + total]
        else:
            out += b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(body) + 25) + body + struct.pack("<II", zlib.crc32(raw), len(raw))
        at += total
    open(dst, "wb").write(bytes(out))


@pytest.mark.parametrize("shape,cfg_key,spec", [("tair10", "configTAIR10", dict(max_nh=20)), ("flybase6", "configFlybase6", dict(max_nh=8, paired=True, rna_seq=True))])
def test_inflate_and_records_on_the_cpu(tool, tmp_path, shape, cfg_key, spec):
    cfg_path = str(tmp_path / "c.txt")
    open(cfg_path, "w").write(CFGS[cfg_key])
    synth = host.Synth(shape, 31, gene_scale=0.05, **spec)
    gtf, bam = str(tmp_path / "a.gtf"), str(tmp_path / "r.bam")
    synth.write_annotation(gtf)
    synth.write_bam_parallel(bam, 0, 40000, 3)
    ann = host.Annotation(host.Config(cfg_path), gtf)
    raw = gzip.open(bam, "rb").read(1 << 22)
    ltext = struct.unpack_from("<I", raw, 4)[0]
    p = 8 + ltext
    nref = struct.unpack_from("<I", raw, p)[0]
    p += 4
    names = []
    for _ in range(nref):
        ln = struct.unpack_from("<I", raw, p)[0]
        names.append(raw[p + 4:p + 4 + ln].split(b"\0")[0].decode())
        p += 8 + ln
    chrs = ann.chromosomes()
    has = set(int(c) for c in np.unique(ann.chr))
    r2c = np.array([chrs.index(nm) if (nm in chrs and chrs.index(nm) in has) else 0xFFFFFF for nm in names], np.uint32)
    files = {"level1": bam}
    for level in (0, 6, 9):
        files["level%d" % level] = str(tmp_path / ("r%d.bam" % level))
        recompress(bam, files["level%d" % level], level)
    for strand, code in (("F", 1), ("R", 2), ("U", 0)):
        hits = host.read_hits(ann, bam, strand)[0]
        dump = str(tmp_path / "hits.bin")
        with open(dump, "wb") as f:
            f.write(struct.pack("<Q", hits.n))
            for k in ("start", "end", "meta", "nh"):
                f.write(np.ascontiguousarray(getattr(hits, k), np.uint32).tobytes())
            f.write(np.ascontiguousarray(hits.read_key, np.uint64).tobytes())
            f.write(r2c.tobytes())
        for name, path in files.items():
            if strand != "F" and name != "level1":
                continue
            pr = subprocess.run([tool, path, dump, str(code)], capture_output=True, text=True)
            assert pr.returncode == 0, (name, strand, pr.stdout[-300:])


@pytest.mark.parametrize("mask", ["0xFull", "0xFFFull"])
def test_read_key_verification_on_the_cpu(tmp_path, mask):
    """bamParseMember (the body of k_bam_parse) with the read keys cut down to 4 / 12 bits: exactly the members that hold a pair
    of neighbouring records with one key and two names -- inside the member or across the border to the next member with a
    record -- must raise BAM_KEY_COLLISION (the tool finds the expected members with one plain walk over all records).  The
    uncut keys raise nothing: that is the `flags` check of the test above."""
    common.ensure_built(("host",))
    exe = build_tool(str(tmp_path / "bam_host_check_keys"), ["-DMMA_NAME_KEY_MASK=" + mask, "-DEXPECT_COLLISIONS"])
    synth = host.Synth("tair10", 32, gene_scale=0.05, max_nh=4)
    bam = str(tmp_path / "r.bam")
    synth.write_bam_parallel(bam, 0, 120000, 3)  # (three parts: members without records lie between them)
    pr = subprocess.run([exe, bam, "-", "1"], capture_output=True, text=True)  # (no expected hits in this mode)
    assert pr.returncode == 0, pr.stdout[-300:]
    words = pr.stdout.split()
    pairs, flagged, border = int(words[3]), int(words[6]), int(words[9])
    assert pairs > 0 and flagged > 0
    if mask == "0xFull":
        assert border > 0, "no colliding pair across a member border in this input: the border check did not run"
