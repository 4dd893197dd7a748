# Same-box A/B of an environment knob: bash scripts/gpu_ab_env.sh VAR valueA valueB [rounds]   (bench.py resident step time, alternating)
VAR=$1; A=$2; B=$3; R=${4:-3}
mkdir -p gpurun_out
for i in $(seq 1 $R); do
  for v in "$A" "$B"; do
    env $VAR=$v timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-side-legs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$VAR=$v', 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'host', d['host_enqueue_ms_per_step'], 'ksum', d['kernel_ms_sum'])"
  done
done | tee gpurun_out/ab_$VAR.log
