mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_golden.py tests/test_gpu_synth_cli.py -m gpu -x -q --timeout 280 > gpurun_out/r2l_tests.txt 2>&1; tail -4 gpurun_out/r2l_tests.txt | cut -c1-200
summ='import json,sys
t=sys.stdin.read().strip()
if not t: print("NO OUTPUT"); sys.exit()
d=json.loads(t); r=d["roofline"]
print("value %.3e ms/step %.2f frac %.3f batch_ms %s"%(d["value"],d["ms_per_step"],r["frac"],r["kernel_ms_per_step"]))'
for w in tair10_srna flybase6_paired; do
timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-file 2>gpurun_out/bench_err.log | python -c "$summ"
done
CMD="python bench.py --workload flybase6_paired --reads 8000000 --steps 1 --warmup 2 --no-cpu-baseline --no-secondary --no-file"
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'^k_batch' -s 4 -c 1 -f -o gpurun_out/prof_r2l_fb $CMD > gpurun_out/ncu_r2l.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_r2l.log
