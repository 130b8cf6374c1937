#!/bin/bash
mkdir -p gpurun_out
ncu --set full --cache-control none --warp-sampling-interval 0 --clock-control none --import-source on -k regex:"lc_" -s 4 -c 4 -o gpurun_out/lc_r2h -f python tools/lt_cut_probe.py vitg fp16 1 > gpurun_out/ncu_lc.log 2>&1
tail -2 gpurun_out/ncu_lc.log
