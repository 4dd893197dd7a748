mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r2l_pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-side-legs --no-cpu-baseline 2>gpurun_out/r2l_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print(d['value'], d['ms_per_step'], d['e2e']['value']); print({x:k[x] for x in k if 'star' in x})"
tail -3 gpurun_out/r2l_bench.err
