#!/bin/bash
# round-2a: cutoff path of the per-layer select -- exactness, timing, tests
mkdir -p gpurun_out
python tools/lt_cut_probe.py vitg fp16 5 > gpurun_out/lt_cut_vitg.log 2>&1
python tools/lt_cut_probe.py t5 bf16 5 > gpurun_out/lt_cut_t5.log 2>&1
python tools/lt_cut_probe.py llama fp16 5 > gpurun_out/lt_cut_llama.log 2>&1
ECF_LT_NSIGMA=0 python tools/lt_cut_probe.py vitg fp16 3 > gpurun_out/lt_cut_vitg_fallback.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "layer_thresh" > gpurun_out/pytest_lt.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lc_|layer_thresh" -c 40 --csv --log-file gpurun_out/lt_cut_launches.csv python tools/lt_cut_probe.py vitg fp16 2 > /dev/null 2>&1
tail -n 12 gpurun_out/lt_cut_vitg.log gpurun_out/lt_cut_t5.log gpurun_out/lt_cut_llama.log gpurun_out/lt_cut_vitg_fallback.log gpurun_out/pytest_lt.log
grep -o '"lc_[a-z_]*[^"]*","[0-9]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[0-9.]*"' gpurun_out/lt_cut_launches.csv | head -0
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/lt_cut_launches.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[1:9]: print(r[ki][:50], r[vi])
PY
