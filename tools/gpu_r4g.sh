#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r4g.log 2>&1
tail -3 gpurun_out/pytest_gpu_r4g.log; echo "tests wall: $SECONDS s"
SECONDS=0
timeout 1500 python bench.py > gpurun_out/bench_r4g.json 2> gpurun_out/bench_r4g.err
echo "bench wall: $SECONDS s"
grep -E "Error|error|Traceback" gpurun_out/bench_r4g.err | tail -5
SECONDS=0
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r4g.json 2> gpurun_out/bench_ref_r4g.err
echo "ref wall: $SECONDS s"; tail -c 200 gpurun_out/bench_ref_r4g.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sqnorm|row_select|lc_|layer_thresh|norm_exchange" -c 2500 --csv --log-file gpurun_out/launches_bench_r4g.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-aten --no-sparsegpt-kernels --prune-wall none > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_bench_r4g.csv
ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:"row_select_tma" -s 4 -c 2 -o gpurun_out/rs_r4g -f env RS_BLOCKS=enc ECF_RS_CORUN=0 python tools/rs_block.py ncu > gpurun_out/ncu_rs.log 2>&1
tail -1 gpurun_out/ncu_rs.log
