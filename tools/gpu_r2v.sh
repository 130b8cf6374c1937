#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/p2p_check.py 2>&1 | grep -v "^n=\|Setting OMP\|\*\*\*" | tee gpurun_out/p2p_check_n2_r2v.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 5 --warmup 3 --prune-wall none --no-sparsegpt-kernels --no-e2e > gpurun_out/bench_n2_r2v.json 2> gpurun_out/bench_n2_r2v.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2_r2v.json').read().strip().splitlines()[-1])
print('N=2 weak ms', d['ms_per_step'], 'value', d['value'], 'strong', d['strong_scaling'])
PY
