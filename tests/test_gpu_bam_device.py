"""BAM decode on the device (mma_submit_bam: BGZF inflate + record parse in CUDA) through the drop-in command line: same table,
counters and warnings as the reference binary and as the host decoder, with chunks cut in many places; files the device route
does not take (XA alternative hits, records across BGZF members) must fall back to the host decoder, loudly at -v level."""
import json
import os
import subprocess

import pytest

from tests import common
from oracle import pyoracle
from mmannot_b200 import host

pytestmark = pytest.mark.gpu

CFGS = json.load(open(os.path.join(common.GOLDEN, "configs.json")))
CLI = os.path.join(common.ROOT, "mmannot_b200", "bin", "mmannot_b200")
SHAPES = [("tair10", "configTAIR10", dict(max_nh=20)), ("hs38", "configHS38", dict(max_nh=100)),
          ("flybase6", "configFlybase6", dict(max_nh=8, paired=True, rna_seq=True))]


def run(cmd, **env):
    return subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, **env))


def report(err):
    return [ln for ln in err.splitlines() if ln.startswith("\t#") or "Warning" in ln or "lines read, done" in ln]


@pytest.mark.parametrize("shape,cfg_key,spec", SHAPES, ids=[s[0] for s in SHAPES])
def test_device_decode_equals_reference_and_host_decode(tmp_path, shape, cfg_key, spec):
    ref_exe = pyoracle.ref_binary("fixed")
    if ref_exe is None:
        pytest.skip("reference binary not built")
    cfg_path = str(tmp_path / (cfg_key + ".txt"))
    open(cfg_path, "w").write(CFGS[cfg_key])
    synth = host.Synth(shape, 991, gene_scale=0.1, **spec)
    gtf, bam = str(tmp_path / "a.gtf"), str(tmp_path / "reads.bam")
    synth.write_annotation(gtf)
    synth.write_bam_parallel(bam, 0, 120000, 3)
    for extra in (["-s", "F"], ["-s", "R", "-y", "ratio"], ["-s", "U", "-l", "10", "-y", "unique"]):
        base = ["-a", gtf, "-c", cfg_path, "-r", bam] + extra
        ref = run([ref_exe] + base)
        dev = run([CLI] + base, MMANNOT_B200_VERBOSE="1")
        small = run([CLI] + base, MMANNOT_B200_VERBOSE="1", MMANNOT_B200_BAM_CHUNK_MB="1", MMANNOT_B200_BAM_LAUNCH_MB="2")
        hst = run([CLI] + base, MMANNOT_B200_HOST_DECODE="1")
        assert dev.returncode == 0, dev.stderr[-800:]
        assert "device BAM decoder not used" not in dev.stderr, dev.stderr[-400:]
        assert "device BAM decoder not used" not in small.stderr
        for got in (dev, small, hst):
            assert got.stdout == ref.stdout
            assert report(got.stderr) == report(ref.stderr)


def test_unknown_chromosomes_warned_in_order(tmp_path):
    """Reads on chromosomes the annotation does not know: the reference's warnings, in order of first appearance."""
    ref_exe = pyoracle.ref_binary("fixed")
    if ref_exe is None:
        pytest.skip("reference binary not built")
    cfg_path = str(tmp_path / "c.txt")
    open(cfg_path, "w").write(CFGS["configTAIR10"])
    gtf, bam = str(tmp_path / "a.gtf"), str(tmp_path / "reads.bam")
    host.Synth("tair10", 5, gene_scale=0.05, max_nh=4).write_annotation(gtf)
    host.Synth("flybase6", 6, gene_scale=0.05, max_nh=4).write_bam(bam, 0, 30000)  # other chromosome names
    base = ["-a", gtf, "-c", cfg_path, "-r", bam, "-s", "F"]
    ref = run([ref_exe] + base)
    dev = run([CLI] + base, MMANNOT_B200_VERBOSE="1")
    assert dev.returncode == 0 and "device BAM decoder not used" not in dev.stderr
    assert dev.stdout == ref.stdout
    assert report(dev.stderr) == report(ref.stderr)
    assert sum("Warning" in ln for ln in report(dev.stderr)) >= 2


def test_files_left_to_the_host_decoder(tmp_path):
    """XA alternative hits (the shipped test BAM) and records across BGZF members: decoded on the host, same results."""
    ref_exe = pyoracle.ref_binary("fixed")
    if ref_exe is None:
        pytest.skip("reference binary not built")
    cfg_path = str(tmp_path / "c.txt")
    open(cfg_path, "w").write(CFGS["configTAIR10"])
    synth = host.Synth("tair10", 17, gene_scale=0.05, max_nh=6)
    gtf, bam = str(tmp_path / "a.gtf"), str(tmp_path / "straddle.bam")
    synth.write_annotation(gtf)
    synth.write_bam(bam, 0, 40000, straddle=True)
    base = ["-a", gtf, "-c", cfg_path, "-r", bam, "-s", "F"]
    ref = run([ref_exe] + base)
    dev = run([CLI] + base, MMANNOT_B200_VERBOSE="1")
    assert dev.returncode == 0
    assert "device BAM decoder not used" in dev.stderr and "across BGZF members" in dev.stderr
    assert dev.stdout == ref.stdout and report(dev.stderr) == report(ref.stderr)
    data = os.path.join(common.ROOT, "tests", "golden")
    ref_bam = "/root/reference/test_dataset.bam"
    if os.path.exists(ref_bam):  # (build container only: the GPU box has no reference tree)
        cfg2 = str(tmp_path / "hs.txt")
        open(cfg2, "w").write(CFGS["configHS38"])
        base = ["-a", "/root/reference/test_dataset.gtf", "-c", cfg2, "-r", ref_bam, "-s", "U"]
        ref = run([ref_exe] + base)
        dev = run([CLI] + base, MMANNOT_B200_VERBOSE="1")
        assert "device BAM decoder not used" in dev.stderr
        assert dev.stdout == ref.stdout
