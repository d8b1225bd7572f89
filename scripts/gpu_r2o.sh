mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bam_device.py -m gpu -q --timeout 280 > gpurun_out/r2o_tests.txt 2>&1; tail -3 gpurun_out/r2o_tests.txt | cut -c1-200
python - > gpurun_out/r2o_file.txt 2>&1 <<'PY'
import os, sys, subprocess, tempfile, time
sys.path.insert(0, ".")
import bench
tmp = tempfile.mkdtemp()
wl = bench.Workload("tair10_srna", tmp)
bam = tmp + "/big.bam"
wl.synth.write_bam_parallel(bam, 0, 25000000, 16); print("%.2f GB" % (os.path.getsize(bam) / 1e9))
cmd = ["mmannot_b200/bin/mmannot_b200", "-a", wl.gtf_path, "-c", wl.config_path, "-s", "F", "-o", os.devnull, "-r", bam]
for rep in range(3):
    t0 = time.time()
    pr = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, MMANNOT_B200_TIMING="1"))
    print("wall %.2fs" % (time.time() - t0), [l for l in pr.stderr.splitlines() if l.startswith("[timing]")])
PY
cat gpurun_out/r2o_file.txt
