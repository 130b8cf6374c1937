#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r4t.log 2>&1
tail -2 gpurun_out/pytest_gpu_r4t.log
timeout 600 python bench.py --no-cpu --no-aten --prune-wall none --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['ms_per_step'], {k:round(v['frac'],3) for k,v in d['roofline']['kernels'].items()}, d['gpu_launches'], d['clocks'])"
