#!/bin/bash
mkdir -p gpurun_out
python tools/rs_block.py 2>&1 | tee gpurun_out/rs_block.log
python tools/kernel_probe.py row_select 2>&1 | grep row_select | tee gpurun_out/kernel_probe_rs.log
timeout 600 python -m pytest tests -x -q -m gpu -k "row" > gpurun_out/pytest_r2p.log 2>&1; tail -3 gpurun_out/pytest_r2p.log
