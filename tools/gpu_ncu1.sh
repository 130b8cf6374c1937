#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
NCU="ncu --clock-control none --set full --import-source on"
timeout 300 $NCU -k regex:layer_thresh -s 1 -c 1 -f -o $O/layer_block_vitg_r1b python tools/one_kernel.py layer_block 0 0 fp16 2 > $O/ncu_lt.log 2>&1
timeout 300 $NCU -k regex:row_select -s 2 -c 1 -f -o $O/row_select_t5enc_r1b python tools/rs_block.py ncu > $O/ncu_rs.log 2>&1
timeout 300 python tools/rs_block.py 2>&1 | tee $O/rs_block.log
ls -la $O/*.ncu-rep
