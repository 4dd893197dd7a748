# First GPU call of the next round (about 2 minutes of box time): is everything still green, and do the opt-in star backward
# variants (tx_star_bwd.cu) pass / pay off?  Outputs under gpurun_out/.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
echo "--- star backward v0 (whole egonets per warp) vs staged"
(TAXO_STAR_BWD_TEST=1 timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -k star_backward 2>&1 | tail -2)
echo "--- star backward, team variant (NOT yet run on hardware when this script was written)"
(TAXO_STAR_BWD_TEST=1 TAXO_STAR_BWD_COOP=24 timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -k star_backward 2>&1 | tail -4)
for cfg in "TAXO_STAR_BWD=0" "TAXO_STAR_BWD=1" "TAXO_STAR_BWD=1 TAXO_STAR_BWD_COOP=24" "TAXO_STAR_BWD=1 TAXO_STAR_BWD_COOP=12" "TAXO_STAR_BWD=1 TAXO_STAR_BWD_COOP=40"; do
  env $cfg timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$cfg', d['value'], d['ms_per_step'], {x:k[x] for x in k if 'bwd' in x and 'gat' in x})"
done
