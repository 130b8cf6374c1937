#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "upop or llama" > gpurun_out/pytest_r2f.log 2>&1
tail -40 gpurun_out/pytest_r2f.log
