#!/bin/bash
mkdir -p gpurun_out
for st in 2 1; do
ECF_ZO_PREFIX_STRIDE=$st timeout 900 python tools/prune_wall.py ecoflap > gpurun_out/prune_wall_eco_r3q_s$st.json 2> gpurun_out/prune_wall_eco_r3q_s$st.err
echo "stride $st"; tail -c 300 gpurun_out/prune_wall_eco_r3q_s$st.json; echo
done
