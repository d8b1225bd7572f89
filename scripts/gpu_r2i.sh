# round 2, session I (1 GPU): device BAM decode tests, file -> table timing, full suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bam_device.py -m gpu -x -q --timeout 280 > gpurun_out/r2i_bam_tests.txt 2>&1; tail -25 gpurun_out/r2i_bam_tests.txt | cut -c1-220
timeout 1500 python -m pytest tests -m gpu -q --timeout 280 --deselect tests/test_gpu_bam_device.py > gpurun_out/r2i_tests.txt 2>&1; tail -6 gpurun_out/r2i_tests.txt | cut -c1-200
timeout 900 python bench.py --no-secondary --no-cpu-baseline --file-reads 25000000 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.log; echo "bench rc=$?"; tail -3 gpurun_out/r2i_bench.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2i_bench.json"))
print("value %.3e frac %.3f e2e %.3e" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"]))
print("e2e_file", d.get("e2e_file"))
PY
