mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -4
summ='import json,sys
t=sys.stdin.read().strip()
if not t: print("NO OUTPUT"); sys.exit()
d=json.loads(t); r=d["roofline"]
print("value %.3e ms/step %.2f frac %.3f batch_ms %s"%(d["value"],d["ms_per_step"],r["frac"],r["kernel_ms_per_step"]))'
for wl in tair10_srna flybase6_paired; do
B="python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-file --no-secondary"
timeout 600 $B 2>gpurun_out/bench_err.log | python -c "$summ"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_batch" -c 120 --csv --log-file gpurun_out/r02u_${wl}.csv $B > gpurun_out/ncu1.log 2>&1
done
