#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r3i.log 2>&1
tail -3 gpurun_out/pytest_gpu_r3i.log
timeout 900 python bench.py --no-cpu --no-aten --prune-wall none > gpurun_out/bench_r3i.json 2> gpurun_out/bench_r3i.err
tail -c 1500 gpurun_out/bench_r3i.json
