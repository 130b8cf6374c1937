#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/prune_wall.py sparsegpt > gpurun_out/prune_wall_sgpt.json 2> gpurun_out/prune_wall_sgpt.err
cat gpurun_out/prune_wall_sgpt.json; grep -i "spent" gpurun_out/prune_wall_sgpt.err | tail -5
