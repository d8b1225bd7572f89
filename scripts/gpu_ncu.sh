# ncu of the dominant kernel on a full-size batch (after the same command has run clean without ncu)
mkdir -p gpurun_out
CMD="python bench.py --reads 8000000 --steps 1 --warmup 1 --no-cpu-baseline $BENCH_ARGS"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'^k_batch$' -s 0 -c 1 -o gpurun_out/prof $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu2.log
