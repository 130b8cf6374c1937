#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "row_select" > gpurun_out/pytest_r3l.log 2>&1
tail -2 gpurun_out/pytest_r3l.log
for c in 0 1; do echo "ECF_RS_CORUN=$c"; ECF_RS_CORUN=$c timeout 300 python tools/rs_block.py; done 2>&1 | tee gpurun_out/rs_block_r3l.log
PROBE_TAG=r3l timeout 600 python tools/kernel_probe.py row_select 2>&1 | tee gpurun_out/kernel_probe_r3l.log
