# quick GPU check: parity tests, then bench variants (no cpu baseline), then one full ncu capture of the dominant kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -8
summ='import json,sys
t=sys.stdin.read().strip()
if not t: print("NO OUTPUT"); sys.exit()
d=json.loads(t); r=d["roofline"]
print("value %.3e e2e %.3e ms/step %.2f frac %.3f batch_ms %s miss %.4f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],r["frac"],r["kernel_ms_per_step"],r["segment_table_miss_frac"]))'
for lib in "" ${VARIANTS:-mmannot_b200/lib/variants/f3.so mmannot_b200/lib/variants/f5.so}; do
  echo "== lib=$lib"
  MMANNOT_B200_LIB=$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BENCH_ARGS 2>gpurun_out/bench_err.log | python -c "$summ"
done
if [ "${LEGACY:-1}" = "1" ]; then
echo "== legacy k_batch"
MMANNOT_B200_LEGACY_BATCH=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BENCH_ARGS 2>/dev/null | python -c "$summ"
fi
if [ "${NCU:-1}" = "1" ]; then
CMD="python bench.py --reads 16000000 --steps 1 --warmup 1 --no-cpu-baseline $BENCH_ARGS"
MMANNOT_B200_LIB=$NCU_LIB timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
MMANNOT_B200_LIB=$NCU_LIB ncu --set full --clock-control none --import-source on -k regex:'^k_batch' -s 0 -c 1 -f -o gpurun_out/prof_q $CMD > gpurun_out/ncu_q.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_q.log
fi
