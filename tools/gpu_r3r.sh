#!/bin/bash
mkdir -p gpurun_out
(for b in 16 24 32 48 64; do echo "band $b"; ECF_RS_BAND=$b timeout 300 python tools/rs_block.py; done) 2>&1 | tee gpurun_out/rs_block_r3r.log
