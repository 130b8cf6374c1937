#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "first_order or global or real_score" > gpurun_out/pytest_r2ad.log 2>&1; grep -v "sparsity:" gpurun_out/pytest_r2ad.log | tail -5
timeout 900 python tools/prune_wall.py ecoflap_first > gpurun_out/prune_wall_first.json 2> gpurun_out/prune_wall_first.err
cat gpurun_out/prune_wall_first.json; grep -i "error\|spent" gpurun_out/prune_wall_first.err | tail -6
