# One GPU session: parity tests, smoke, bench, then the ncu launch list and one full capture of the top kernel.
mkdir -p gpurun_out
set -x
nvidia-smi -L; nproc
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -15
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.log; echo "bench rc=$?"; tail -5 gpurun_out/bench_a.log; cat gpurun_out/bench_a.json
if [ "${NCU:-1}" = "1" ]; then
timeout 300 python bench.py --reads 8000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_a.csv python bench.py --reads 8000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_batch' -s 4 -c 4 -o gpurun_out/prof_a python bench.py --reads 8000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
echo "ncu rc=$?"
fi
