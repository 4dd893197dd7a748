for mode in off off inline inline thread; do
TAXO_SAMPLER=$mode timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/bq.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e']['value'], d['e2e']['host_ms_per_step']['fwd_bwd'], d['clocks'])"
done
tail -2 gpurun_out/bq.err
