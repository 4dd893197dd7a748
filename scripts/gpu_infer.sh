mkdir -p gpurun_out
timeout 200 python scripts/infer_bench.py 2>&1 | tail -4
TAXO_STAR_FWD=0 timeout 200 python scripts/infer_bench.py 2>&1 | tail -4
TAXO_STAR_CHUNK=16 timeout 200 python scripts/infer_bench.py 2>&1 | tail -4
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e']['value'], d['clocks'])"
