ls -la oracle/_ref oracle/_ref/model 2>&1 | head
timeout 300 python scripts/host_prof.py 2>&1 | tail -70
