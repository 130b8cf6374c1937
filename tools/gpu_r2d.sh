#!/bin/bash
# round-2d: baseline for this round -- full GPU tests, bench line, launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r2d.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2d.log
timeout 900 python bench.py > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err
tail -c 3000 gpurun_out/bench_r2d.json
