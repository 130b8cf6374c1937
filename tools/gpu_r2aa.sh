#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python bench.py > gpurun_out/bench_r2aa.json 2> gpurun_out/bench_r2aa.err
echo "bench wall: $SECONDS s"
grep -E "Error|error|Traceback" gpurun_out/bench_r2aa.err | tail -5
SECONDS=0
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2aa.json 2> gpurun_out/bench_ref_r2aa.err
echo "ref wall: $SECONDS s"; tail -c 400 gpurun_out/bench_ref_r2aa.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_bench_r2aa.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-aten --no-sparsegpt-kernels --prune-wall none > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_bench_r2aa.csv
