mkdir -p gpurun_out
for b in 4194304 8388608 16777216 33554432 67108864; do
  echo "== batch=$b"
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --batch $b 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3e e2e %.3e ms/step %.2f frac %.3f batch_ms %s launches %d'%(d['value'],d['e2e']['value'],d['ms_per_step'],r['frac'],r['kernel_ms_per_step'],d['gpu_launches']))"
done
for fs in 5 7 8; do
  echo "== fast shift=$fs"
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --fast-shift $fs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3e ms/step %.2f frac %.3f batch_ms %s miss %.4f index %d'%(d['value'],d['ms_per_step'],r['frac'],r['kernel_ms_per_step'],r['segment_table_miss_frac'],d['config']['index_bytes']))"
done
