"""File -> table wall clock of the drop-in command line against the reference binary on the same synthetic BAM/GFF
(GPU box).  Writes gpurun_out/cli_compare.json."""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import common
from oracle import pyoracle
from mmannot_b200 import host

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
CFGS = json.load(open(os.path.join(common.GOLDEN, "configs.json")))
tmp = tempfile.mkdtemp(prefix="cli_cmp_")
cfg = os.path.join(tmp, "configTAIR10.txt"); open(cfg, "w").write(CFGS["configTAIR10"])
synth = host.Synth("tair10", 20261020, max_nh=20)
gtf = os.path.join(tmp, "a.gff"); synth.write_annotation(gtf)
bam = os.path.join(tmp, "reads.bam"); synth.write_bam(bam, 0, reads)
records = synth.count_hits(0, reads)
args = ["-a", gtf, "-c", cfg, "-r", bam, "-s", "F"]
out = {"reads": reads, "records": int(records), "bam_mb": os.path.getsize(bam) / 1e6, "host_cores": os.cpu_count()}
cli = os.path.join(ROOT, "mmannot_b200", "bin", "mmannot_b200")
ref = pyoracle.ref_binary("fixed")
res = {}
for name, exe in (("b200_cli", cli), ("reference", ref)):
    best = None
    for rep in range(2):
        t0 = time.perf_counter()
        pr = subprocess.run([exe] + args, capture_output=True, text=True)
        dt = time.perf_counter() - t0
        assert pr.returncode == 0, pr.stderr[-500:]
        best = dt if best is None else min(best, dt)
    res[name] = pr.stdout
    out[name + "_wall_s"] = best
    out[name + "_records_per_s"] = records / best
assert res["b200_cli"] == res["reference"], "tables differ"
out["tables_identical"] = True
out["speedup"] = out["reference_wall_s"] / out["b200_cli_wall_s"]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "cli_compare.json"), "w"), indent=1)
print(json.dumps(out))
