#!/bin/bash
mkdir -p gpurun_out
for g in 16 32; do echo "== group $g"; ECF_BLOCK_GRAPH_GROUP=$g timeout 900 python tools/prof_sweep.py sparsegpt 2>/dev/null | grep -v "^$" | head -40; done > gpurun_out/prof_sweep_r3y.log 2>&1
cut -c1-160 gpurun_out/prof_sweep_r3y.log
