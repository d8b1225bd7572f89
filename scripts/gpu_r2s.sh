mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-file --no-secondary"
timeout 400 $B > gpurun_out/plain1.log 2>&1 && tail -c 600 gpurun_out/plain1.log && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02s_launches.csv $B > gpurun_out/ncu1.log 2>&1
echo "ncu list rc=$?"
B2="python bench.py --reads 16000000 --steps 1 --warmup 1 --no-cpu-baseline --no-file --no-secondary --device-batch 33554432"
ncu --set full --clock-control none --import-source on -k regex:"^k_batch_lean" -s 0 -c 1 -f -o gpurun_out/r02s_prof $B2 > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"^k_batch_walk" -s 0 -c 1 -f -o gpurun_out/r02s_walk $B2 > gpurun_out/ncu3.log 2>&1
echo "ncu walk rc=$?"
