# scaling run on one box: N = $1 ranks
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --no-side-legs --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
fi
tail -2 gpurun_out/scale_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/scale_n$N.json').read().strip().splitlines()[-1])
print('N=$N', 'value', d['value'], 'ms', d['ms_per_step'], 'host', d['host_enqueue_ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['host_ms_per_step'], 'rows-e2e', d['e2e']['feature_rows_variant']['value'], 'prop_ro', d.get('propagate_readout'), d['host'])
PY
