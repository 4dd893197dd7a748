# Round 2, call A: GPU tests green?  star backward variants correct / fast?  ncu --set full of the star backward (v0) at L0 + L1.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) | tee gpurun_out/r2a_pytest.log
echo "--- star backward v0"
(TAXO_STAR_BWD_TEST=1 timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -k star_backward 2>&1 | tail -2)
echo "--- star backward team"
(TAXO_STAR_BWD_TEST=1 TAXO_STAR_BWD_COOP=24 timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -k star_backward 2>&1 | tail -4)
for cfg in "TAXO_STAR_BWD=0" "TAXO_STAR_BWD=1" "TAXO_STAR_BWD=1 TAXO_STAR_BWD_COOP=24" "TAXO_STAR_BWD=1 TAXO_STAR_BWD_COOP=12" "TAXO_STAR_BWD=1 TAXO_STAR_BWD_COOP=40"; do
  env $cfg timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$cfg', d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], {x:k[x] for x in k if 'bwd' in x and 'gat' in x})"
done
TAXO_STAR_BWD=1 timeout 600 ncu -k regex:'gat_star_bwd' --launch-skip 6 -c 2 --set full --import-source on --clock-control none -f -o gpurun_out/r2a_star_bwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu.log 2>&1
tail -2 gpurun_out/r2a_ncu.log
ls -la gpurun_out/r2a_star_bwd.ncu-rep
