#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pruners.py -q -m gpu > gpurun_out/pytest_r4m.log 2>&1
grep -n "passed\|failed\|^E \|FAILED" gpurun_out/pytest_r4m.log | head -20
echo "memo on : $(timeout 900 python tools/prune_wall.py wanda sparsegpt 2>/dev/null | tail -c 300)"
echo "memo off: $(ECF_TOWER_MEMO=0 timeout 900 python tools/prune_wall.py wanda sparsegpt 2>/dev/null | tail -c 300)"
