#!/bin/bash
mkdir -p gpurun_out
for g in 4 8 16 32; do echo "sparsegpt group $g: $(ECF_BLOCK_GRAPH_GROUP=$g timeout 900 python tools/prune_wall.py sparsegpt 2>/dev/null | tail -c 120)"; done
echo "sparsegpt eager: $(ECF_BLOCK_GRAPH=0 timeout 900 python tools/prune_wall.py sparsegpt 2>/dev/null | tail -c 120)"
echo "wanda: $(timeout 900 python tools/prune_wall.py wanda 2>/dev/null | tail -c 120)"
