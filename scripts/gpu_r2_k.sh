mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gather_rows or plan" 2>&1 | tail -3
timeout 600 python bench.py --steps 30 --warmup 5 --no-side-legs --no-cpu-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
tail -5 gpurun_out/r2k_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2k_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','host_enqueue_ms_per_step','kernel_ms_sum','star_bwd_reruns')}); print(d['e2e'])
PY
