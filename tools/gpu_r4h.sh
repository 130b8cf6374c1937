#!/bin/bash
mkdir -p gpurun_out
export RS_BLOCKS=enc ECF_RS_CORUN=0
ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:"row_select_tma" -s 2 -c 2 -o gpurun_out/rs_r4h -f python tools/rs_block.py ncu > gpurun_out/ncu_rs.log 2>&1
tail -1 gpurun_out/ncu_rs.log
unset ECF_RS_CORUN RS_BLOCKS
python tools/rs_block.py 2>&1 | tee gpurun_out/rs_block_r4h.log
