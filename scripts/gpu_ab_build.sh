# Same-box A/B of a COMPILE-TIME knob: bash scripts/gpu_ab_build.sh "<nvcc flags A>" "<nvcc flags B>" [rounds]
# (rebuilds the library between runs; prints the resident step time and the kernels whose name matches $KERNELS)
A=$1; B=$2; R=${3:-2}; K=${KERNELS:-readout}
mkdir -p gpurun_out
for i in $(seq 1 $R); do
  for v in "$A" "$B"; do
    TAXO_NVCC_FLAGS="$v" python -m taxoexpan_b200.build --force > /dev/null 2>&1
    timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-side-legs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']
print('[$v]', 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], {a:b for a,b in k.items() if '$K' in a})"
  done
done | tee gpurun_out/ab_build.log
python -m taxoexpan_b200.build --force > /dev/null 2>&1
