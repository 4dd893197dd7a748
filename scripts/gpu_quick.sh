# Quick GPU check: GPU tests + one bench line (no CPU baseline leg, no side legs). Outputs under gpurun_out/.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-side-legs > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
tail -n 3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','host_enqueue_ms_per_step','kernel_ms_sum')}, d['e2e']['value'], d['e2e']['ms_per_step'])
print({k:v for k,v in d['kernel_ms_per_step'].items() if 'star' in k or 'gemm' in k})
print(d['roofline'])
PY
