bash scripts/gpu_round2.sh
set +x
summ='import json,sys
t=sys.stdin.read().strip()
if not t: print("NO OUTPUT"); sys.exit()
d=json.loads(t); r=d["roofline"]
print("value %.3e ms/step %.2f frac %.3f %s batch_ms %s"%(d["value"],d["ms_per_step"],r["frac"],r["kernel"],r["kernel_ms_per_step"]))'
echo "== hs38 default / bins"
timeout 600 python bench.py --workload hs38_multi --steps 5 --warmup 3 --no-cpu-baseline --no-file --no-secondary 2>gpurun_out/hs38_a.log | python -c "$summ"
MMANNOT_B200_MAX_BINS=60000000 timeout 600 python bench.py --workload hs38_multi --steps 5 --warmup 3 --no-cpu-baseline --no-file --no-secondary 2>gpurun_out/hs38_b.log | python -c "$summ"
tail -3 gpurun_out/hs38_b.log
echo "== coordsorted launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"slow|Radix|iota|gather|k_batch_lean" -c 2000 --csv --log-file gpurun_out/r02v_coord.csv python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-file --secondary-reads 2000000 > gpurun_out/ncu_coord.log 2>&1
echo "rc=$?"
