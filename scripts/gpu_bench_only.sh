mkdir -p gpurun_out
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_n1.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e']['value'], d['clocks'], d['cpu_baseline']['value'])"
