#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_r4q.json 2> gpurun_out/bench_n2_r4q.err
echo "bench N=2 wall: $SECONDS s rc=$?"
grep -E "Error|error|Traceback" gpurun_out/bench_n2_r4q.err | tail; tail -2 gpurun_out/bench_n2_r4q.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2_r4q.json').read().strip().splitlines()[-1])
print('N=2 weak ms', d['ms_per_step'], 'value', d['value'])
print('strong', d['strong_scaling'])
print('walls', {k:(v['wall_s'], v['sparsity']) for k,v in d['prune_wall_s'].items()})
PY
