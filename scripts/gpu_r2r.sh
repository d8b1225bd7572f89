mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
summ='import json,sys
t=sys.stdin.read().strip()
if not t: print("NO OUTPUT"); sys.exit()
d=json.loads(t); r=d["roofline"]
print("value %.3e ms/step %.2f frac %.3f batch_ms %s"%(d["value"],d["ms_per_step"],r["frac"],r["kernel_ms_per_step"]))'
for lib in "" mmannot_b200/lib/variants/t640r96.so; do
echo "== $lib"
MMANNOT_B200_LIB=$lib timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-file 2>gpurun_out/bench_err.log | python -c "$summ"
done
