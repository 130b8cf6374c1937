#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "row_select" > gpurun_out/pytest_r4p.log 2>&1
tail -2 gpurun_out/pytest_r4p.log
python tools/rs_block.py 2>&1 | tee gpurun_out/rs_block_r4p.log
ECF_RS_CORUN=0 python tools/rs_block.py 2>&1 | tee -a gpurun_out/rs_block_r4p.log
