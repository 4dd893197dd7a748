mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r2j_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-side-legs > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
tail -5 gpurun_out/r2j_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','host_enqueue_ms_per_step','kernel_ms_sum','star_bwd_reruns')}, d['e2e'])
print(d['cpu_baseline'])
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-400
