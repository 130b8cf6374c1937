#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pruners.py -q -m gpu -k "zeroth or ecoflap or stage1 or first_order" > gpurun_out/pytest_r4d.log 2>&1
grep -n "passed\|failed\|^E \|FAILED" gpurun_out/pytest_r4d.log | head -20
echo "ecoflap static : $(timeout 900 python tools/prune_wall.py ecoflap 2>/dev/null | tail -c 200)"
echo "ecoflap per-batch: $(ECF_ZO_STATIC=0 timeout 900 python tools/prune_wall.py ecoflap 2>/dev/null | tail -c 200)"
echo "ecoflap static stride 2: $(ECF_ZO_PREFIX_STRIDE=2 timeout 900 python tools/prune_wall.py ecoflap 2>/dev/null | tail -c 200)"
