# Round 2, call B: the new parity tests (full-size golden, general-CSR stress, per-row f16x3 error).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "full_baseline_size or stress or heavy_tailed or per_row" -s 2>&1 | tail -40 | tee gpurun_out/r2b_pytest.log
