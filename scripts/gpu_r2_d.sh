mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "star_backward" 2>&1 | tail -8
timeout 600 ncu -k regex:'gat_star_bwd' --launch-skip 12 -c 4 --set full --import-source on --clock-control none -f -o gpurun_out/r2d_star_bwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_ncu.log 2>&1
tail -2 gpurun_out/r2d_ncu.log
