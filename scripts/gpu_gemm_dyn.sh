mkdir -p gpurun_out
for v in 0 16 32 0 16 32; do
  echo "== TAXO_GEMM_DBG=$v"
  TAXO_GEMM_DBG=$v TAXO_GEMM_SLAB_ROWS=32 timeout 300 python scripts/gemm_bench.py 2>&1 | sed -e 's/tf32x3 .* ms (.* TF\/s eff)   f16x3/f16x3/' | grep "NT fwd\|NT dz L1"
done | tee gpurun_out/gemm_dyn2.log
