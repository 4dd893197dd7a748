# Round-end check: GPU tests, smoke(), both bench arms (N = 1) -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total,driver_version --format=csv > gpurun_out/smi.txt
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3) > gpurun_out/pytest_gpu.log
tail -n 1 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_n1.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e']['value'], d['clocks'], d['cpu_baseline']['value'], d['roofline']['frac'], [(r['kernel'], round(r['frac'],3)) for r in d['roofline_all']])"
