#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "bulk_copy" > gpurun_out/pytest_r3c.log 2>&1
tail -3 gpurun_out/pytest_r3c.log
ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:"row_select_tma" -s 2 -c 2 -o gpurun_out/rs_r3c -f python tools/rs_block.py ncu > gpurun_out/ncu_rs.log 2>&1
tail -2 gpurun_out/ncu_rs.log
