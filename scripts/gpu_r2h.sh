# round 2, session H (2 GPUs): CLI -G diff
mkdir -p gpurun_out
python - > gpurun_out/r2h_cli_diff.txt 2>&1 <<'PY'
import json, os, subprocess, sys, tempfile
sys.path.insert(0, ".")
from mmannot_b200 import host
from oracle import pyoracle
cfgs = json.load(open("tests/golden/configs.json"))
tmp = tempfile.mkdtemp()
cfg = tmp + "/c.txt"; open(cfg, "w").write(cfgs["configTAIR10"])
synth = host.Synth("tair10", 4242, gene_scale=0.1, max_nh=20)
gtf, bam = tmp + "/a.gtf", tmp + "/reads.bam"
synth.write_annotation(gtf); synth.write_bam(bam, 0, 60000)
base = ["-a", gtf, "-c", cfg, "-r", bam, "-s", "F"]
ref = subprocess.run([pyoracle.ref_binary("fixed")] + base, capture_output=True, text=True)
for g in ("1", "2"):
    got = subprocess.run(["mmannot_b200/bin/mmannot_b200"] + base + ["-G", g], capture_output=True, text=True)
    print("G", g, "rc", got.returncode, "table equal", got.stdout == ref.stdout)
    a, b = got.stdout.splitlines(), ref.stdout.splitlines()
    print(len(a), len(b))
    sa, sb = set(a), set(b)
    for l in sorted(sa - sb)[:8]: print("  only ours:", repr(l))
    for l in sorted(sb - sa)[:8]: print("  only ref :", repr(l))
PY
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_fuzz.py -m gpu -q --timeout 280 -k "multi or export_import or allreduce or torchrun" > gpurun_out/r2h_tests.txt 2>&1; tail -15 gpurun_out/r2h_tests.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --no-file --no-cpu-baseline > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.log; echo "n2 rc=$?"; grep -i "error" gpurun_out/r2h_bench_n2.log | head -5
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2h_bench_n2.json"))
print("N=2 value %.3e ms %.3f frac %.3f e2e %.3e" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]))
for k,v in d["workloads"].items(): print("  ", k, "%.3e" % v["value"], "ms %.2f" % v["ms_per_step"], "frac %.3f" % v["roofline"]["frac"], v["roofline"]["kernel"], v["roofline"]["kernel_ms_per_step"], len(v["checks"]))
PY
