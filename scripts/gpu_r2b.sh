# round 2, session A: parity of the new k_batch_lean, bench of the block-shape variants, full ncu capture
mkdir -p gpurun_out
nvidia-smi -L; nproc
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -8
summ='import json,sys
t=sys.stdin.read().strip()
if not t: print("NO OUTPUT"); sys.exit()
d=json.loads(t); r=d["roofline"]
print("value %.3e e2e %.3e ms/step %.2f frac %.3f batch_ms %s miss %.4f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],r["frac"],r["kernel_ms_per_step"],r["segment_table_miss_frac"]))'
for lib in "" ${VARIANTS}; do
  echo "== lib=$lib"
  MMANNOT_B200_LIB=$lib timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BENCH_ARGS 2>gpurun_out/bench_err.log | tee gpurun_out/r2c_bench_$(basename "$lib" .so).json | python -c "$summ"
done
echo "== no bins (k_batch_fast)"
MMANNOT_B200_NO_BINS=1 timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BENCH_ARGS 2>/dev/null | python -c "$summ"
echo "== flybase6_paired"
timeout 400 python bench.py --workload flybase6_paired --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err_fb.log | tee gpurun_out/r2c_bench_flybase.json | python -c "$summ"
CMD="python bench.py --reads 16000000 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'^k_batch' -s 0 -c 1 -f -o gpurun_out/prof_r2c $CMD > gpurun_out/ncu_r2c.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_r2c.log
