mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
summ='import json,sys
t=sys.stdin.read().strip()
if not t: print("NO OUTPUT"); sys.exit()
d=json.loads(t); r=d["roofline"]
print("value %.3e ms/step %.2f frac %.3f batch_ms %s"%(d["value"],d["ms_per_step"],r["frac"],r["kernel_ms_per_step"]))
for k,v in d.get("workloads",{}).items(): print(k, "%.3e" % v["value"], "ms %.2f" % v["ms_per_step"], "frac %.3f" % v["roofline"]["frac"], v["roofline"]["kernel"], v["roofline"]["kernel_ms_per_step"])'
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-file 2>gpurun_out/bench_err.log | python -c "$summ"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-file --no-secondary"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_batch" -c 120 --csv --log-file gpurun_out/r02t_launches.csv $B > gpurun_out/ncu1.log 2>&1
echo "ncu list rc=$?"
