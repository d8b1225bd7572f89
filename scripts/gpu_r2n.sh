mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_synth_cli.py -m gpu -q --timeout 280 -k "random" > gpurun_out/r2n_tests.txt 2>&1; tail -3 gpurun_out/r2n_tests.txt | cut -c1-200
python - > gpurun_out/r2n_prep.txt 2>&1 <<'PY'
import os, sys, tempfile
sys.path.insert(0, ".")
import bench
tmp = "/tmp/r2n"; os.makedirs(tmp, exist_ok=True)
wl = bench.Workload("tair10_srna", tmp)
wl.synth.write_bam_parallel(tmp + "/big.bam", 0, 8000000, 16)
open(tmp + "/cmd.txt", "w").write(" ".join(["mmannot_b200/bin/mmannot_b200", "-a", wl.gtf_path, "-c", wl.config_path, "-s", "F", "-o", "/dev/null", "-r", tmp + "/big.bam"]))
PY
CMD=$(cat /tmp/r2n/cmd.txt)
MMANNOT_B200_TIMING=1 $CMD 2>&1 | grep timing
ncu --set full --clock-control none --import-source on -k regex:'k_bam_inflate' -c 1 -f -o gpurun_out/prof_r2n_inflate $CMD > gpurun_out/ncu_r2n.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_r2n.log
