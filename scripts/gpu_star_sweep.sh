mkdir -p gpurun_out
for cfg in "TAXO_STAR_PREFETCH=2" "TAXO_STAR_PREFETCH=1" "TAXO_STAR_PREFETCH=0" "TAXO_STAR_PREFETCH=2 TAXO_STAR_CHUNK=8" "TAXO_STAR_PREFETCH=1 TAXO_STAR_CHUNK=8" "TAXO_STAR_PREFETCH=0 TAXO_STAR_CHUNK=8"; do
  env $cfg timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$cfg', d['ms_per_step'], {x:k[x] for x in k if 'fwd' in x and 'gat' in x})"
done
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
