# round 2, session F (2 GPUs): whole GPU suite incl. multi-GPU tests, bench at N=1 (default line) and N=2 (torchrun)
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 280 2>&1 | tail -6
timeout 900 python bench.py --no-file > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.log; echo "n1 rc=$?"; tail -3 gpurun_out/r2f_bench_n1.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --no-file > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.log; echo "n2 rc=$?"; tail -5 gpurun_out/r2f_bench_n2.log
python - <<'PY'
import json
for n in (1, 2):
    try:
        d=json.load(open("gpurun_out/r2f_bench_n%d.json" % n))
    except Exception as e:
        print("N", n, "no line:", e); continue
    print("N=%d value %.3e ms %.3f frac %.3f e2e %.3e" % (n, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]))
    for k,v in d["workloads"].items(): print("  ", k, "%.3e" % v["value"], "ms %.2f" % v["ms_per_step"], "frac %.3f" % v["roofline"]["frac"], v["roofline"]["kernel"], v["roofline"]["kernel_ms_per_step"], len(v["checks"]))
    print("  ", d["checks"])
PY
