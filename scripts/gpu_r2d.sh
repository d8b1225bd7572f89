# round 2, session D: carveout / block shape sweep of k_batch_lean (bench only, no ncu)
mkdir -p gpurun_out
summ='import json,sys
t=sys.stdin.read().strip()
if not t: print("NO OUTPUT"); sys.exit()
d=json.loads(t); r=d["roofline"]
print("value %.3e e2e %.3e ms/step %.2f frac %.3f batch_ms %s miss %.4f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],r["frac"],r["kernel_ms_per_step"],r["segment_table_miss_frac"]))'
timeout 600 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_golden.py -m gpu -x -q --timeout 180 2>&1 | tail -3
for cv in "" 70 86 100; do
  echo "== carveout=$cv"
  MMANNOT_B200_CARVEOUT=$cv timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err.log | tee gpurun_out/r2d_bench_cv$cv.json | python -c "$summ"
done
for lib in ${VARIANTS}; do
  for cv in "" 100; do
  echo "== lib=$lib carveout=$cv"
  MMANNOT_B200_CARVEOUT=$cv MMANNOT_B200_LIB=$lib timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err.log | python -c "$summ"
  done
done
echo "== flybase6_paired"
timeout 400 python bench.py --workload flybase6_paired --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err_fb.log | tee gpurun_out/r2d_bench_flybase.json | python -c "$summ"
echo "== hs38"
timeout 400 python bench.py --workload hs38_multi --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err_hs.log | tee gpurun_out/r2d_bench_hs38.json | python -c "$summ"
