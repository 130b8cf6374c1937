#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_r2ab.sh
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sqnorm|row_select|lc_|layer_thresh|norm_exchange" -c 2500 --csv --log-file gpurun_out/launches_bench_r2aa.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-aten --no-sparsegpt-kernels --prune-wall none > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_bench_r2aa.csv
