# One-shot evidence run on the GPU box (round 2): GPU tests, bench (both arms), the cuBLAS-backend yardstick, the ncu launch list of the
# bench command, ncu --set full of the star kernels and the GEMMs.  Outputs land in gpurun_out/ (merged back);
# scripts/make_profiles.py turns them into the tracked files under profiles/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total,driver_version --format=csv > gpurun_out/smi.txt
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3) > gpurun_out/pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
TAXO_GEMM=cublas timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-side-legs > gpurun_out/bench_cublas.json 2> gpurun_out/bench_cublas.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side-legs > gpurun_out/b_ncu.log 2>&1
timeout 700 ncu -k regex:'gat_star_bwd|gat_star_fwd|gemm_tf32x3' --launch-skip 30 -c 30 --set full --import-source on --clock-control none -f -o gpurun_out/kernels_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-side-legs > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/pytest_gpu.log
tail -n 2 gpurun_out/ncu_full.log
head -c 600 gpurun_out/bench_n1.json
