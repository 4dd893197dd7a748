mkdir -p gpurun_out
timeout 600 ncu -k regex:'gat_star_fwd' --launch-skip 6 -c 2 --set full --import-source on --clock-control none -f -o gpurun_out/r2m_star_fwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-side-legs > gpurun_out/r2m_ncu.log 2>&1
tail -2 gpurun_out/r2m_ncu.log
