#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pruners.py -q -m gpu -k "replay or sparsegpt" > gpurun_out/pytest_r3v.log 2>&1
grep -n "passed\|failed\|^E \|FAILED\|PASSED" gpurun_out/pytest_r3v.log | head -30
