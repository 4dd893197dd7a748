set -x
mkdir -p gpurun_out
for R in 8 12 24; do
  TAXO_BWD2_TILE_ROWS=$R timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('R=$R', d['ms_per_step'], {x:k[x] for x in k if 'staged' in x})"
done
timeout 600 ncu -k regex:gat_bwd_staged --launch-skip 6 -c 2 --set full --import-source on --clock-control none -f -o gpurun_out/bwd_staged python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bwd.log 2>&1
tail -3 gpurun_out/ncu_bwd.log
mkdir -p gpurun_out; ls -la gpurun_out/
