#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pruners.py -x -q -m gpu -k "zeroth or ecoflap or stage" > gpurun_out/pytest_r3p.log 2>&1
tail -15 gpurun_out/pytest_r3p.log
timeout 900 python tools/prune_wall.py ecoflap > gpurun_out/prune_wall_eco_r3p.json 2> gpurun_out/prune_wall_eco_r3p.err
tail -c 500 gpurun_out/prune_wall_eco_r3p.json; grep -n "spent\|captured\|Error" gpurun_out/prune_wall_eco_r3p.err | tail -8
