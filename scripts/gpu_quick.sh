# quick GPU check: fuzz + golden tests, then bench variants (no cpu baseline)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -8
for lib in "" mmannot_b200/lib/variants/b2.so mmannot_b200/lib/variants/b4.so; do
  echo "== lib=$lib"
  MMANNOT_B200_LIB=$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BENCH_ARGS 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3e e2e %.3e ms/step %.2f frac %.3f batch_ms %s miss %.4f'%(d['value'],d['e2e']['value'],d['ms_per_step'],r['frac'],r['kernel_ms_per_step'],r['segment_table_miss_frac']))"
done
