# One GPU session for the round's evidence: parity tests, smoke, the bench line of both arms, then the ncu launch list of the bench
# command and one full capture of the top kernel (each only after its command has exited 0 without ncu).
mkdir -p gpurun_out
set -x
nvidia-smi -L; nproc
timeout 1500 python -m pytest tests -m gpu -q --timeout 280 > gpurun_out/r02_tests.txt 2>&1; tail -4 gpurun_out/r02_tests.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.log; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.log; echo "ref rc=$?"; cat gpurun_out/r02_bench_reference.json | cut -c1-400
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-file --no-secondary > gpurun_out/plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-file --no-secondary > gpurun_out/ncu1.log 2>&1
echo "ncu list rc=$?"
timeout 300 python bench.py --reads 16000000 --steps 1 --warmup 1 --no-cpu-baseline --no-file --no-secondary --device-batch 33554432 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"^k_batch" -s 0 -c 1 -f -o gpurun_out/r02_prof python bench.py --reads 16000000 --steps 1 --warmup 1 --no-cpu-baseline --no-file --no-secondary --device-batch 33554432 > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench.json"))
print("value %.3e frac %.3f e2e %.3e" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"]))
print("e2e_file", d.get("e2e_file"))
print("cpu", d.get("cpu_baseline"))
for k,v in d["workloads"].items(): print(k, "%.3e" % v["value"], "ms %.2f" % v["ms_per_step"], "frac %.3f" % v["roofline"]["frac"], v["roofline"]["kernel"], v["roofline"]["kernel_ms_per_step"])
PY
