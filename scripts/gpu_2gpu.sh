# final 2-GPU check: multi-GPU tests, N=2 bench (both with and without the secondary workloads would be too long: no secondary)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 280 > gpurun_out/r2z_tests.txt 2>&1; tail -5 gpurun_out/r2z_tests.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --no-file --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.log; echo "n2 rc=$?"; grep -i "error" gpurun_out/r02_bench_n2.log | head -5
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_n2.json"))
print("N=2 value %.3e ms %.3f frac %.3f e2e %.3e" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]))
print(d["checks"])
PY
