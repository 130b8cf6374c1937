#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r4l.log 2>&1
tail -2 gpurun_out/pytest_gpu_r4l.log; echo "tests wall: $SECONDS s"
SECONDS=0
timeout 1500 python bench.py > gpurun_out/bench_r4l.json 2> gpurun_out/bench_r4l.err
echo "bench wall: $SECONDS s"; grep -E "Error|error|Traceback" gpurun_out/bench_r4l.err | tail -5
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r4l.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], {k:round(v['frac'],3) for k,v in d['roofline']['kernels'].items()}, {k:v['wall_s'] for k,v in d['prune_wall_s'].items()}, d['e2e']['ms_per_step'])
PY
