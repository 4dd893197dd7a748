TAXO_NVCC_FLAGS=-DTX_BWD_PROFILE python -m taxoexpan_b200.build --force > /dev/null 2>&1
timeout 200 python scripts/bwd_prof.py 2>&1 | tail -12
