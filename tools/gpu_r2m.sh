#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python bench.py > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err
echo "bench wall: $SECONDS s"
grep -E "Error|error|Traceback" gpurun_out/bench_r2m.err | tail; tail -5 gpurun_out/bench_r2m.err
tail -c 300 gpurun_out/bench_r2m.json
