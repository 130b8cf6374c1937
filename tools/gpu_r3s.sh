#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r3s.log 2>&1
tail -3 gpurun_out/pytest_gpu_r3s.log; echo "tests wall: $SECONDS s"
SECONDS=0
timeout 1500 python bench.py > gpurun_out/bench_r3s.json 2> gpurun_out/bench_r3s.err
echo "bench wall: $SECONDS s"
grep -E "Error|error|Traceback" gpurun_out/bench_r3s.err | tail -5
SECONDS=0
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r3s.json 2> gpurun_out/bench_ref_r3s.err
echo "ref wall: $SECONDS s"; tail -c 300 gpurun_out/bench_ref_r3s.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_bench_r3s.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-aten --no-sparsegpt-kernels --prune-wall none > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_bench_r3s.csv
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
