#!/bin/bash
mkdir -p gpurun_out
ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:"row_select_tma" -s 2 -c 2 -o gpurun_out/rs_r3f -f python tools/rs_block.py ncu > gpurun_out/ncu_rs.log 2>&1
tail -2 gpurun_out/ncu_rs.log
