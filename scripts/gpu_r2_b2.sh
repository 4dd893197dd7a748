mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "full_baseline_size or stress or heavy_tailed or per_row" 2>&1 | tail -30 | tee gpurun_out/r2b_pytest.log
timeout 600 python scripts/dbg_full_dh.py 2>&1 | tail -30 | tee gpurun_out/r2b_dbg.log
