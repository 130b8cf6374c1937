#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "row_select" > gpurun_out/pytest_r3b.log 2>&1
tail -5 gpurun_out/pytest_r3b.log
for st in 0 1 2; do echo "ECF_RS_TMA=$st"; ECF_RS_TMA=$st timeout 300 python tools/rs_block.py; done 2>&1 | tee gpurun_out/rs_block_r3b.log
