#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pruners.py tests/test_host_rows_cpu.py -q > gpurun_out/pytest_r4s.log 2>&1
tail -2 gpurun_out/pytest_r4s.log
