# bench every tuning build under mmannot_b200/lib/variants (kernel time per step and roofline fraction)
summ='import json,sys
t=sys.stdin.read().strip()
if not t: print("NO OUTPUT"); sys.exit()
d=json.loads(t); r=d["roofline"]
print("value %.3e ms/step %.2f frac %.3f batch_ms %.3f"%(d["value"],d["ms_per_step"],r["frac"],r["kernel_ms_per_step"]["ms_batch"]))'
for lib in "" mmannot_b200/lib/variants/*.so; do
  echo "== lib=$lib"
  MMANNOT_B200_LIB=$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-wide $BENCH_ARGS 2>gpurun_out/bench_err.log | python -c "$summ"
done
