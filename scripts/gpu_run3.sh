mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for R in 0 8 12 24; do
  TAXO_BWD2_TILE_ROWS=$R timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('R=$R', d['ms_per_step'], {x:k[x] for x in k if 'staged' in x})"
done
