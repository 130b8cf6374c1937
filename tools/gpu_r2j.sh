#!/bin/bash
mkdir -p gpurun_out
for pdl in 0 1; do
echo "== PDL=$pdl"
ECF_LT_PDL=$pdl python tools/lt_cut_probe.py vitg fp16 5 2>&1 | grep -v "thres ok" | tee gpurun_out/lt_cut_vitg_pdl$pdl.log
ECF_LT_PDL=$pdl python tools/lt_cut_probe.py llama fp16 3 2>&1 | grep -v "thres ok" | tee gpurun_out/lt_cut_llama_pdl$pdl.log
done
timeout 900 python -m pytest tests -x -q -m gpu -k "layer_thresh" > gpurun_out/pytest_r2j.log 2>&1
tail -5 gpurun_out/pytest_r2j.log
