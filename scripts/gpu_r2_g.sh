mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/r2g_pytest.log
for cfg in "TAXO_LAYER_CALL=1" "TAXO_LAYER_CALL=0"; do
  env $cfg timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2g_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$cfg', d['value'], d['ms_per_step'], 'host', d['host_enqueue_ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('host_ms_per_step'), 'launches', d['gpu_launches'], 'ksum', d['kernel_ms_sum']); print({x:k[x] for x in k})"
done
tail -3 gpurun_out/r2g_bench.err
