#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pruners.py -q -m gpu > gpurun_out/pytest_r3z.log 2>&1
grep -n "passed\|failed\|^E \|FAILED" gpurun_out/pytest_r3z.log | head -20
for g in 16 32; do echo "sparsegpt group $g: $(ECF_BLOCK_GRAPH_GROUP=$g timeout 900 python tools/prune_wall.py sparsegpt 2>/dev/null | tail -c 120)"; done
echo "sparsegpt eager: $(ECF_BLOCK_GRAPH=0 timeout 900 python tools/prune_wall.py sparsegpt 2>/dev/null | tail -c 120)"
echo "wanda eager: $(timeout 900 python tools/prune_wall.py wanda 2>/dev/null | tail -c 120)"
echo "wanda g8: $(ECF_BLOCK_GRAPH_GROUP=8 timeout 900 python tools/prune_wall.py wanda 2>/dev/null | tail -c 120)"
echo "ecoflap: $(timeout 900 python tools/prune_wall.py ecoflap 2>/dev/null | tail -c 200)"
