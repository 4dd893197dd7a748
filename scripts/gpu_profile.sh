# One-shot evidence run: GPU tests, bench (both arms), ncu launch list of the bench command, ncu --set full of the fused kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > gpurun_out/smi.txt
(timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -3) > gpurun_out/pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu -k regex:'gat_bwd_staged|gat_fused_fwd' --launch-skip 8 -c 4 --set full --import-source on --clock-control none -f -o gpurun_out/fused_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/pytest_gpu.log gpurun_out/ncu_full.log
cat gpurun_out/bench_n1.json | head -c 1500
