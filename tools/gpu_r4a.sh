#!/bin/bash
mkdir -p gpurun_out
echo "ecoflap flush-free: $(timeout 900 python tools/prune_wall.py ecoflap 2>/dev/null | tail -c 200)"
echo "ecoflap torch ctx : $(ECF_GRAPH_CAPTURE_FLUSH=1 timeout 900 python tools/prune_wall.py ecoflap 2>/dev/null | tail -c 200)"
echo "ecoflap flush-free stride 6: $(ECF_ZO_PREFIX_STRIDE=6 timeout 900 python tools/prune_wall.py ecoflap 2>/dev/null | tail -c 200)"
