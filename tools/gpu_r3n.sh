#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "bulk_copy or batched" > gpurun_out/pytest_r3n.log 2>&1
tail -2 gpurun_out/pytest_r3n.log
(echo "ECF_RS_CORUN=0"; ECF_RS_CORUN=0 timeout 300 python tools/rs_block.py
for sh in 2 3 4; do for pad in 62 76 100; do echo "CORUN short=$sh pad=$pad"; ECF_RS_CORUN_SHORT=$sh ECF_RS_CORUN_PAD_KB=$pad timeout 300 python tools/rs_block.py; done; done) 2>&1 | tee gpurun_out/rs_block_r3n.log
