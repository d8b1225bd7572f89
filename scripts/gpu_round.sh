# One GPU session: parity tests, smoke, bench, then the ncu launch list of the bench command and one full capture of the top kernel.
mkdir -p gpurun_out
set -x
nvidia-smi -L; nproc
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.log; echo "bench rc=$?"; tail -3 gpurun_out/bench_a.log; cat gpurun_out/bench_a.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.log; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
if [ "${NCU:-1}" = "1" ]; then
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_a.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
echo "ncu list rc=$?"
timeout 300 python bench.py --reads 16000000 --steps 1 --warmup 1 --no-cpu-baseline --device-batch 33554432 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"^k_batch" -s 0 -c 1 -o gpurun_out/prof_a python bench.py --reads 16000000 --steps 1 --warmup 1 --no-cpu-baseline --device-batch 33554432 > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"
fi
timeout 600 python tests/tools/cli_compare.py 5000000 2>&1 | tail -1
