# NG ranks (default 4): does the overlapped gradient all-reduce slow the kernels it runs beside?  NCCL CTA budget / bucket variants
mkdir -p gpurun_out; NG=${NG:-4}
run() {
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $NG --steps 40 --warmup 8 --no-cpu-baseline --no-side-legs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']
print('$*', 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'prop_ro', d['propagate_readout']['ms_per_step'], 'star_bwd_L0', k.get('tx_gat_star_bwd[L0]'), 'dz_L0', k.get('gemm_dz[L0]'), 'dw_L0', k.get('gemm_dw[L0]'))"
}
for r in 1; do
IFS=";" read -ra VS <<< "${VARIANTS:-TAXO_BUCKET_SEGMENTS=1;TAXO_BUCKET_SEGMENTS=2;TAXO_BUCKET_SEGMENTS=2 TAXO_BUCKET_GATE=0;TAXO_BUCKET_SEGMENTS=3}"; for v in "${VS[@]}"; do run $v; done
done 2>&1 | tee gpurun_out/n2_nccl.log
