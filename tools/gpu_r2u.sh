#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r2u.log 2>&1
grep -v "sparsity:" gpurun_out/pytest_gpu_r2u.log | tail -12
