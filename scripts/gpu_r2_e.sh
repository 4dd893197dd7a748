mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "star_backward" 2>&1 | grep -E "passed|failed|Error|assert|err " | head -20
for cfg in "TAXO_STAR_CHUNK=4" "TAXO_STAR_CHUNK=8" "TAXO_STAR_CHUNK=16"; do
  env $cfg timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2e_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$cfg', d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e']['value']); print({x:k[x] for x in k if 'star' in x})"
done
tail -3 gpurun_out/r2e_bench.err
timeout 600 ncu -k regex:'gat_star_bwd' --launch-skip 12 -c 4 --set full --import-source on --clock-control none -f -o gpurun_out/r2e_star_bwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_ncu.log 2>&1
tail -2 gpurun_out/r2e_ncu.log
