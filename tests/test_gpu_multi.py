"""Several GPUs on one sample (needs >= 2 devices; skipped otherwise): reads dealt out by read name, the tables summed on the
devices.  Three routes, each against the oracle / the reference on the UNSHARDED input:
  * mma_allreduce  -- the contexts of one process, NCCL clique inside the library (what the command line's -G uses)
  * multi.merge_on_device under torch.distributed.run -- one process per GPU (what bench.py times)
  * the command line with -G 2 against the reference binary's table and counters"""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import common, fuzz
from oracle import pyoracle
from mmannot_b200 import host

pytestmark = pytest.mark.gpu


def _n_gpus():
    from mmannot_b200 import device
    return int(device.lib().mma_device_count())


needs2 = pytest.mark.skipif("_n_gpus() < 2", reason="needs two GPUs")


@needs2
@pytest.mark.parametrize("strategy,overlap,shuffle", [("default", -1.0, False), ("unique", 1.0, False), ("ratio", -1.0, False), ("default", 0.5, True)])
def test_allreduce_in_one_process(strategy, overlap, shuffle):
    import torch  # noqa: F401  (first: a process that loads torch later must not already hold another libnccl.so.2)
    from mmannot_b200 import device, multi
    n = min(_n_gpus(), 4)
    rng = np.random.default_rng(77)
    et = fuzz.make_elements(rng)
    feats = fuzz.make_features(rng, et, n_feat=500)
    hits = fuzz.make_hits(rng, feats, n_reads=30000, max_nh=9, messy=0.2)
    if shuffle:
        hits = fuzz.shuffle_hits(rng, hits, block=100)
    ref = pyoracle.run(et.elem_line, et.elem_strand, et.elem_vicinity, feats, hits, strategy=strategy, overlap=overlap)
    shard = multi.shard_of_keys(hits.read_key, n)
    anns = [device.Annotator(et, strategy=strategy, overlap=overlap, max_batch_hits=9973, device=g) for g in range(n)]
    try:
        for g, a in enumerate(anns):
            a.load_features(feats)
            sel = shard == g
            a.submit(0, host.Hits(*[np.ascontiguousarray(getattr(hits, k)[sel]) for k in ("start", "end", "meta", "nh", "read_key")]))
        device.allreduce(anns, 0)
        for a in anns:  # the merged result is on every GPU
            res = a.finish(0)
            got = device.values_by_mask(res["rows"])
            assert set(got) == set(ref["rows"])
            for m, v in ref["rows"].items():
                assert abs(got[m] - v) <= 1e-9 * max(1.0, abs(v))
            assert res["stats"] == ref["stats"]
    finally:
        for a in anns:
            a.close()


@needs2
def test_merge_on_device_under_torchrun():
    env = dict(os.environ, PYTHONPATH=common.ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29641",
           os.path.join(common.ROOT, "tests", "tools", "merge_worker.py")]
    pr = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=280)
    assert pr.returncode == 0, pr.stderr[-2000:]
    assert pr.stdout.count(" ok") == 2


@needs2
@pytest.mark.parametrize("shape,cfg_key,spec", [("tair10", "configTAIR10", dict(max_nh=20)), ("flybase6", "configFlybase6", dict(max_nh=8, paired=True, rna_seq=True))])
def test_command_line_spread_over_two_gpus(tmp_path, shape, cfg_key, spec):
    import json
    cfgs = json.load(open(os.path.join(common.GOLDEN, "configs.json")))
    cfg_path = str(tmp_path / (cfg_key + ".txt"))
    open(cfg_path, "w").write(cfgs[cfg_key])
    synth = host.Synth(shape, 4242, gene_scale=0.1, **spec)
    gtf, bam = str(tmp_path / "a.gtf"), str(tmp_path / "reads.bam")
    synth.write_annotation(gtf)
    synth.write_bam(bam, 0, 60000)
    ref_exe = pyoracle.ref_binary("fixed")
    if ref_exe is None:
        pytest.skip("reference binary not built")
    cli = os.path.join(common.ROOT, "mmannot_b200", "bin", "mmannot_b200")
    base = ["-a", gtf, "-c", cfg_path, "-r", bam, "-s", "F"]
    ref = subprocess.run([ref_exe] + base, capture_output=True, text=True)
    got = subprocess.run([cli] + base + ["-G", "2"], capture_output=True, text=True)
    assert got.returncode == 0, got.stderr[-1000:]
    assert got.stdout == ref.stdout
    pick = lambda err: [ln for ln in err.splitlines() if ln.startswith("\t#")]
    assert pick(got.stderr) == pick(ref.stderr)
