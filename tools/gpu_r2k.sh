#!/bin/bash
mkdir -p gpurun_out
python tools/rs_block.py > gpurun_out/rs_block.log 2>&1; cat gpurun_out/rs_block.log
timeout 900 python -m pytest tests -x -q -m gpu -k "row_select or row or t5" > gpurun_out/pytest_r2k.log 2>&1
tail -5 gpurun_out/pytest_r2k.log
