"""bench.py --impl reference runs without a GPU (it times the compiled reference, oracle/_ref): the JSON line of the contract,
on a tiny sample.  The other arm needs a B200; what can be checked here is that both arms print the same `config` object."""
import json
import os
import subprocess
import sys

import pytest

from tests import common

REQUIRED = ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "cpu_baseline", "e2e")


def test_reference_arm_prints_the_contract_line():
    from oracle import pyoracle
    if not os.path.exists(pyoracle.ref_binary("fixed")):
        pytest.skip("oracle/_ref not built")
    common.ensure_built(("host",))
    cmd = [sys.executable, os.path.join(common.ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
           "--ref-reads", "5000", "--ref-threads", "2"]
    pr = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=common.ROOT)
    assert pr.returncode == 0, pr.stderr[-500:]
    lines = [ln for ln in pr.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "alignment_records_per_sec" and d["unit"] == "records/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 2
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # the same config object as the b200 arm prints for the same flags
    sys.path.insert(0, common.ROOT)
    import bench
    import argparse
    wl = argparse.Namespace(w=bench.WORKLOADS["tair10_srna"], name="tair10_srna")
    assert d["config"] == bench.bench_config(wl, argparse.Namespace(reads=0), 1)


def test_reference_arm_under_torchrun_prints_one_line_from_rank_0():
    """N > 1 (the driver launches both arms through torch.distributed.run): rank 0 alone times the reference and prints the line,
    the other ranks leave with exit code 0 and print nothing."""
    from oracle import pyoracle
    if not os.path.exists(pyoracle.ref_binary("fixed")):
        pytest.skip("oracle/_ref not built")
    common.ensure_built(("host",))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(common.ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "1", "--ref-reads", "5000", "--ref-threads", "2"]
    pr = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=common.ROOT)
    assert pr.returncode == 0, pr.stderr[-500:]
    lines = [ln for ln in pr.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    sys.path.insert(0, common.ROOT)
    import bench
    import argparse
    wl = argparse.Namespace(w=bench.WORKLOADS["tair10_srna"], name="tair10_srna")
    assert d["config"] == bench.bench_config(wl, argparse.Namespace(reads=0), 2)
