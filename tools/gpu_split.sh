#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
ECF_LT_SPLIT=1 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pruners.py -m gpu -x -q -k "layer or vit or lavis or blip" 2>&1 | tail -3
ECF_LT_SPLIT=1 timeout 100 python tools/one_kernel.py layer_block 0 0 fp16 4 2>&1 | grep "block select" | tail -2
ECF_LT_SPLIT=1 timeout 200 ncu --clock-control none --metrics gpu__time_duration.sum,launch__registers_per_thread,launch__grid_size,smsp__inst_executed.sum -k regex:"lt_split|layer_thresh" -s 5 -c 5 --csv --log-file $O/lt_split_launches.csv python tools/one_kernel.py layer_block 0 0 fp16 3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/lt_split_launches.csv") if l.startswith('"'))]
h=rows[0]; ni=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); ii=h.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((r[ii], r[ni][:28]), {})[r[mi].split("__")[-1][:18]]=r[vi]
for k,v in d.items(): print(k, v)
PY
for sp in 1 0; do ECF_LT_SPLIT=$sp timeout 250 python bench.py --no-e2e --no-cpu > $O/bench_sp$sp.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/bench_sp$sp.json').read().strip().splitlines()[-1]); k=d['roofline']['kernels']['layer_thresh']; print('split=$sp', d['ms_per_step'], k['avg_us'], k['frac'])"; done
