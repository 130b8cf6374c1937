#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8_r4f.json 2> gpurun_out/bench_n8_r4f.err
echo "bench N=8 wall: $SECONDS s rc=$?"
grep -E "Error|error|Traceback" gpurun_out/bench_n8_r4f.err | tail -5; tail -3 gpurun_out/bench_n8_r4f.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n8_r4f.json').read().strip().splitlines()[-1])
print('N=8 weak ms', d['ms_per_step'], 'value', d['value'])
print('strong', d['strong_scaling'])
print('e2e', d['e2e'])
print('walls', {k:(v['wall_s'], v['sparsity']) for k,v in d['prune_wall_s'].items()})
print(d['config']['parallelism'], d['cpu_affinity'])
PY
