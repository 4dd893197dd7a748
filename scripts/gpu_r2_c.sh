# Round 2, call C: new star backward (correctness vs staged, then the whole GPU suite, then bench)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "star_backward" 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r2c_pytest.log
for cfg in "TAXO_STAR_BWD=1" "TAXO_STAR_BWD=0"; do
  env $cfg timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2c_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$cfg', d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e']['value']); print({x:k[x] for x in k})"
done
tail -3 gpurun_out/r2c_bench.err
