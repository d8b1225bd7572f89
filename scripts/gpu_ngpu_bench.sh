mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --no-file --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.log; echo "n$N rc=$?"; grep -i "error" gpurun_out/r02_bench_n$N.log | head -5
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.load(open("gpurun_out/r02_bench_n%s.json"%N))
print("N=%s value %.3e ms %.3f frac %.3f e2e %.3e" % (N, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]))
print(d["checks"])
PY
