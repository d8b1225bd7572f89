mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bam_device.py tests/test_gpu_fuzz.py tests/test_gpu_synth_cli.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.log; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02w_bench.json"))
print("value %.3e frac %.3f e2e %.3e" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"]))
print("e2e_file", d.get("e2e_file"))
for k,v in d["workloads"].items(): print(k, "%.3e" % v["value"], "ms %.2f" % v["ms_per_step"], "frac %.3f" % v["roofline"]["frac"], v["roofline"]["kernel"], v["roofline"]["kernel_ms_per_step"])
PY
