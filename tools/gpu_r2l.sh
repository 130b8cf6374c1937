#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/prune_wall.py wanda ecoflap sparsegpt > gpurun_out/prune_wall.json 2> gpurun_out/prune_wall.err
tail -c 1500 gpurun_out/prune_wall.json; grep "spent" gpurun_out/prune_wall.err | tail -20; tail -5 gpurun_out/prune_wall.err
