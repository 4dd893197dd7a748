# 2 GPUs: NCCL gradient parity test + the 2-rank bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r2_dist_gpu_pytest.log
cat gpurun_out/dist_gpu_parity.json 2>/dev/null
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -2 gpurun_out/r2_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e']['value'], d.get('propagate_readout'))"
