#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "zeroth or ecoflap or blip2 or stage1 or llama" > gpurun_out/pytest_r2s.log 2>&1; grep -v "sparsity:" gpurun_out/pytest_r2s.log | tail -6
timeout 1200 python tools/prune_wall.py ecoflap > gpurun_out/prune_wall_graph.json 2> gpurun_out/prune_wall_graph.err
cat gpurun_out/prune_wall_graph.json; grep -i "captured\|eagerly\|spent" gpurun_out/prune_wall_graph.err | tail -8
