#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pruners.py -q -m gpu > gpurun_out/pytest_r3w.log 2>&1
grep -n "passed\|failed\|^E \|FAILED" gpurun_out/pytest_r3w.log | head -30
timeout 900 python tools/prune_wall.py sparsegpt > gpurun_out/prune_wall_r3w.json 2> gpurun_out/prune_wall_r3w.err
tail -c 300 gpurun_out/prune_wall_r3w.json; grep -n "captured\|Error\|spent" gpurun_out/prune_wall_r3w.err | tail -6
