mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bam_device.py tests/test_gpu_synth_cli.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.log; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02x_bench.json"))
print("value %.3e frac %.3f e2e %.3e" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"]))
print("e2e_file", {k:v for k,v in d.get("e2e_file").items() if k!="what"})
PY
