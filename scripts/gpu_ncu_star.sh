mkdir -p gpurun_out
timeout 600 ncu -k regex:'gat_star_fwd' --launch-skip 8 -c 2 --set full --import-source on --clock-control none -f -o gpurun_out/star_fwd3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_star3.log 2>&1
ls -la gpurun_out/star_fwd3.ncu-rep
