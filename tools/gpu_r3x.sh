#!/bin/bash
mkdir -p gpurun_out
for g in 16 8 4; do
ECF_BLOCK_GRAPH_GROUP=$g timeout 900 python tools/prune_wall.py wanda > gpurun_out/prune_wall_r3x_g$g.json 2> gpurun_out/prune_wall_r3x_g$g.err
echo "group $g: $(tail -c 200 gpurun_out/prune_wall_r3x_g$g.json)"; grep -n "spent" gpurun_out/prune_wall_r3x_g$g.err | tail -4 | tr '\n' ' '; echo
done
ECF_BLOCK_GRAPH_GROUP=32 timeout 900 python tools/prune_wall.py sparsegpt 2>/dev/null | tail -c 200
