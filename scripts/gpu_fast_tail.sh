# validates the chunk-border tail of k_batch_fast (position-map path): fuzz suite without bin entries, hs38 shape against the oracle, hs38 bench
mkdir -p gpurun_out
MMANNOT_B200_NO_BINS=1 timeout 100 python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3
timeout 60 python -m pytest tests/test_gpu_synth_cli.py -m gpu -x -q -k "benchmark_shapes and hs38" 2>&1 | tail -3
timeout 60 python bench.py --workload hs38_multi --steps 5 --warmup 3 --no-cpu-baseline --no-file --no-secondary 2>gpurun_out/hs38_t.log | python -c '
import json,sys
t=sys.stdin.read().strip()
d=json.loads(t); r=d["roofline"]
print("value %.3e ms/step %.2f frac %.3f %s batch_ms %s"%(d["value"],d["ms_per_step"],r["frac"],r["kernel"],r["kernel_ms_per_step"]))'
