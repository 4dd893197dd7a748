mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r2f_pytest.log
for cfg in "TAXO_STAR_BWD_CHUNK=8" "TAXO_STAR_BWD_CHUNK=12" "TAXO_STAR_BWD_CHUNK=6"; do
  env $cfg timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2f_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$cfg', d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e']['value']); print({x:k[x] for x in k if 'star' in x})"
done
tail -3 gpurun_out/r2f_bench.err
