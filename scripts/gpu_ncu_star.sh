mkdir -p gpurun_out
timeout 600 ncu -k regex:'gat_star_fwd|gat_fused_fwd' --launch-skip 8 -c 2 --set full --import-source on --clock-control none -f -o gpurun_out/star_fwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_star.log 2>&1
tail -n 3 gpurun_out/ncu_star.log
ls -la gpurun_out/star_fwd.ncu-rep
