"""Command-line surface of the drop-in (mmannot.cpp:1946-2091) on the paths that end before any GPU work: same exit code
and same first lines on stderr as the reference binary (usage, version, wrong / missing parameters, bad values, missing
configuration, configuration that does not match the annotation)."""
import json
import os
import subprocess

import pytest

from tests import common
from oracle import pyoracle
from mmannot_b200 import host

CLI = os.path.join(common.ROOT, "mmannot_b200", "bin", "mmannot_b200")
REF = pyoracle.ref_binary("fixed")
CFGS = json.load(open(os.path.join(common.GOLDEN, "configs.json")))

pytestmark = pytest.mark.skipif(REF is None or not os.path.exists(CLI), reason="needs oracle/_ref and the CLI binary")


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("cliargs")
    cfg = str(tmp / "configTAIR10.txt"); open(cfg, "w").write(CFGS["configTAIR10"])
    other = str(tmp / "configHS38.txt"); open(other, "w").write(CFGS["configHS38"])
    synth = host.Synth("tair10", 99, gene_scale=0.01, max_nh=4)
    gtf = str(tmp / "a.gff"); synth.write_annotation(gtf)
    bam = str(tmp / "r.bam"); synth.write_bam(bam, 0, 200)
    return dict(cfg=cfg, other=other, gtf=gtf, bam=bam, tmp=str(tmp))


def both(args, cwd):
    out = []
    for exe in (REF, CLI):
        pr = subprocess.run([exe] + args, capture_output=True, text=True, timeout=120, cwd=cwd)
        out.append((pr.returncode, pr.stdout, pr.stderr.split("\n")))
    return out


CASES = [
    ([], 3), (["-h"], 3), (["-v"], 1), (["-z"], 2), (["-a", "{gtf}"], 2), (["-r", "{bam}"], 2),
    (["-a", "{gtf}", "-r", "{bam}", "-n", "a", "b"], 2),
    (["-a", "{gtf}", "-r", "{bam}", "{bam}", "-m", "x.txt"], 2),
    (["-a", "{gtf}", "-r", "{bam}", "{bam}", "-M", "x.txt"], 2),
    (["-a", "{gtf}", "-r", "{bam}", "-c", "/nonexistent/config.txt"], 1),
    (["-a", "{gtf}", "-r", "{bam}", "-c", "{cfg}", "-s", "X"], 2),
    (["-a", "{gtf}", "-r", "{bam}", "-c", "{cfg}", "-s", "FR"], 2),
    (["-a", "{gtf}", "-r", "{bam}", "-c", "{cfg}", "-y", "foo"], 2),
    (["-a", "{gtf}", "-r", "{bam}", "-c", "{cfg}", "-f", "cram"], 2),
    (["-a", "/nonexistent/a.gtf", "-r", "{bam}", "-c", "{cfg}"], 0),
    (["-a", "{gtf}", "-r", "{bam}", "-c", "{other}"], 0),  # configuration does not match the annotation: nothing parsed
]


@pytest.mark.parametrize("args,n_lines", CASES, ids=lambda a: " ".join(a) if isinstance(a, list) else str(a))
def test_same_exit_code_and_messages(files, args, n_lines):
    args = [a.format(**files) for a in args]
    (rc_ref, out_ref, err_ref), (rc, out, err) = both(args, files["tmp"])
    assert rc == rc_ref
    assert out == out_ref
    assert err[:n_lines] == err_ref[:n_lines]
    if n_lines == 0:  # runs that get as far as the annotation: compare everything but the usage text, which lists -g
        assert rc != 0
        assert [l for l in err if l.strip()] [-1:] == [l for l in err_ref if l.strip()][-1:]
