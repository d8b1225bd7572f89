# round 2, session E: dynamic shared-memory layout of k_batch_lean; carveout sweep; ncu capture
mkdir -p gpurun_out
summ='import json,sys
t=sys.stdin.read().strip()
if not t: print("NO OUTPUT"); sys.exit()
d=json.loads(t); r=d["roofline"]
print("value %.3e e2e %.3e ms/step %.2f frac %.3f batch_ms %s miss %.4f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],r["frac"],r["kernel_ms_per_step"],r["segment_table_miss_frac"]))'
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -3
for cv in "" 50 57 64 72 86; do
  echo "== carveout=$cv"
  MMANNOT_B200_CARVEOUT=$cv timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --no-file 2>gpurun_out/bench_err.log | tee gpurun_out/r2e_bench_cv$cv.json | python -c "$summ"
done
for lib in ${VARIANTS}; do
  echo "== lib=$lib"
  MMANNOT_B200_LIB=$lib timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --no-file 2>gpurun_out/bench_err.log | python -c "$summ"
done
echo "== flybase6_paired"
timeout 400 python bench.py --workload flybase6_paired --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --no-file 2>gpurun_out/bench_err_fb.log | tee gpurun_out/r2e_bench_flybase.json | python -c "$summ"
CMD="python bench.py --reads 16000000 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --no-file"
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'^k_batch' -s 0 -c 1 -f -o gpurun_out/prof_r2e $CMD > gpurun_out/ncu_r2e.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_r2e.log
echo "== full default bench line"
timeout 900 python bench.py > gpurun_out/r2e_bench_full.json 2> gpurun_out/r2e_bench_full.log; echo "rc=$?"; tail -5 gpurun_out/r2e_bench_full.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2e_bench_full.json"))
print("value %.3e frac %.3f e2e %.3e" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"]))
print("e2e_file", d.get("e2e_file"))
print("cpu", d.get("cpu_baseline"))
for k,v in d["workloads"].items(): print(k, "%.3e" % v["value"], "ms %.2f" % v["ms_per_step"], "frac %.3f" % v["roofline"]["frac"], v["roofline"]["kernel_ms_per_step"], v["checks"])
print(d["checks"])
PY
