# compute-sanitizer over a representative slice of the GPU tests (full-size cases excluded: the tools slow kernels down 10-100x)
mkdir -p gpurun_out
SEL='(golden and small and (egonet_batch or general_kernels or dgl_batch)) or star_backward or plan_tables or gather_rows or degenerate or native_layer or dropout_matches or batched_general'
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool" | tee -a gpurun_out/r2_sanitizer.log
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "exit code $?" | tee -a gpurun_out/r2_sanitizer.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/r2_sanitizer_$tool.log | tail -5 | tee -a gpurun_out/r2_sanitizer.log
done
