#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/prune_wall.py wanda sparsegpt > gpurun_out/prune_wall_r4i.json 2> gpurun_out/prune_wall_r4i.err
tail -c 400 gpurun_out/prune_wall_r4i.json; grep -n "Error\|Traceback" gpurun_out/prune_wall_r4i.err | head -5
