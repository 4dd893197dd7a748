mkdir -p gpurun_out
nproc; python -c "import os; print(len(os.sched_getaffinity(0)))"
for extra in "" "--no-pin-cores"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 $extra > gpurun_out/bench_n2$extra.json 2> gpurun_out/n2.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_n2$extra.json') if l.startswith('{')][-1]); print('$extra', d['n_gpus'], d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e']['value'], d['e2e']['host_ms_per_step'], d['host'], d['clocks'])"
done
