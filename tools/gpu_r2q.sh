#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_r2q.json 2> gpurun_out/bench_n2_r2q.err
echo "bench N=2 wall: $SECONDS s rc=$?"
grep -E "Error|error|Traceback" gpurun_out/bench_n2_r2q.err | tail; tail -3 gpurun_out/bench_n2_r2q.err
tail -c 1500 gpurun_out/bench_n2_r2q.json
SECONDS=0
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2_r2q.json 2> gpurun_out/bench_ref_n2_r2q.err
echo "ref N=2 wall: $SECONDS s"; tail -c 600 gpurun_out/bench_ref_n2_r2q.json
