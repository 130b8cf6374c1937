#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_pruners.py -x -q -m gpu > gpurun_out/pytest_r3u.log 2>&1
tail -15 gpurun_out/pytest_r3u.log
timeout 900 python tools/prune_wall.py wanda sparsegpt > gpurun_out/prune_wall_r3u.json 2> gpurun_out/prune_wall_r3u.err
tail -c 500 gpurun_out/prune_wall_r3u.json; grep -n "captured\|Error" gpurun_out/prune_wall_r3u.err | tail -5
