#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "bulk_copy" > gpurun_out/pytest_r3g.log 2>&1
tail -2 gpurun_out/pytest_r3g.log
for st in 1 2; do echo "ECF_RS_TMA=$st"; ECF_RS_TMA=$st timeout 300 python tools/rs_block.py; done 2>&1 | tee gpurun_out/rs_block_r3g.log
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"row_select" -s 2 -c 2 python tools/rs_block.py ncu 2>&1 | grep -E "row_select|duration|inst_executed|issue_active" | tee gpurun_out/ncu_rs_r3g.log
