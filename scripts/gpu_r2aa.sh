mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -3
MMANNOT_B200_NO_BINS=1 timeout 600 python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3
summ='import json,sys
t=sys.stdin.read().strip()
if not t: print("NO OUTPUT"); sys.exit()
d=json.loads(t); r=d["roofline"]
print("value %.3e ms/step %.2f frac %.3f %s batch_ms %s"%(d["value"],d["ms_per_step"],r["frac"],r["kernel"],r["kernel_ms_per_step"]))'
timeout 600 python bench.py --workload hs38_multi --steps 5 --warmup 3 --no-cpu-baseline --no-file --no-secondary 2>gpurun_out/hs38_a.log | python -c "$summ"
MMANNOT_B200_NO_BINS=1 timeout 600 python bench.py --workload flybase6_paired --steps 5 --warmup 3 --no-cpu-baseline --no-file --no-secondary 2>gpurun_out/fly_nb.log | python -c "$summ"
