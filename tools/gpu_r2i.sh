#!/bin/bash
mkdir -p gpurun_out
python tools/lt_cut_probe.py vitg fp16 5 > gpurun_out/lt_cut_vitg.log 2>&1
grep -v "thres ok" gpurun_out/lt_cut_vitg.log
python tools/lt_cut_probe.py t5 bf16 3 2>&1 | grep -v "thres ok" > gpurun_out/lt_cut_t5.log; cat gpurun_out/lt_cut_t5.log
python tools/lt_cut_probe.py llama fp16 3 2>&1 | grep -v "thres ok" > gpurun_out/lt_cut_llama.log; cat gpurun_out/lt_cut_llama.log
timeout 900 python -m pytest tests -x -q -m gpu -k "layer_thresh" > gpurun_out/pytest_r2i.log 2>&1
tail -5 gpurun_out/pytest_r2i.log
