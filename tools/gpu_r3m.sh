#!/bin/bash
mkdir -p gpurun_out
for c in 0 1; do echo "ECF_RS_CORUN=$c"; ECF_RS_CORUN=$c timeout 300 python tools/rs_block.py; done 2>&1 | tee gpurun_out/rs_block_r3m.log
