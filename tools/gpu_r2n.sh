#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "global or real_score or get_mask" > gpurun_out/pytest_r2n.log 2>&1
tail -40 gpurun_out/pytest_r2n.log
