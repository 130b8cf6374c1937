#!/bin/bash
# Round-1 GPU pass: parity tests, bench (both arms), ncu launch list, ncu --set full of the hot kernels.
set -x
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi > $O/nvsmi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 600 $O/bench.json
timeout 400 python bench.py --impl reference --steps 1 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
PROBE_TAG=r1a timeout 600 python tools/kernel_probe.py > $O/kernel_probe.log 2>&1
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum -k regex:'sqnorm|row_select|layer_thresh' -c 2500 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/bench_under_ncu.log 2>&1
FULL="$NCU --set full --import-source on"
timeout 300 $FULL -k regex:row_select -s 1 -c 1 -o $O/row_select_2048x2048_bf16 python tools/one_kernel.py row_select 2048 2048 bf16 2 > $O/ncu1.log 2>&1
timeout 300 $FULL -k regex:row_select -s 1 -c 1 -o $O/row_select_5120x2048_bf16 python tools/one_kernel.py row_select 5120 2048 bf16 2 > $O/ncu2.log 2>&1
timeout 300 $FULL -k regex:row_select -s 1 -c 1 -o $O/row_select_2048x5120_bf16 python tools/one_kernel.py row_select 2048 5120 bf16 2 > $O/ncu3.log 2>&1
timeout 300 $FULL -k regex:layer_thresh -s 1 -c 1 -o $O/layer_block_vitg python tools/one_kernel.py layer_block 0 0 fp16 2 > $O/ncu4.log 2>&1
timeout 300 $FULL -k regex:sqnorm -s 2 -c 1 -o $O/sqnorm_vit python tools/one_kernel.py sqnorm_vit 0 0 fp16 1 > $O/ncu5.log 2>&1
timeout 300 $FULL -k regex:sqnorm -s 2 -c 1 -o $O/sqnorm_t5 python tools/one_kernel.py sqnorm_t5 0 0 bf16 1 > $O/ncu6.log 2>&1
timeout 300 $FULL -k regex:hessian -s 1 -c 2 -o $O/hessian_25216x3072_fp16 python tools/one_kernel.py hessian 25216 3072 fp16 3 > $O/ncu7.log 2>&1
timeout 300 $FULL -k regex:obs -c 12 -o $O/obs_3072x768 python tools/one_kernel.py obs 3072 768 fp16 1 > $O/ncu8.log 2>&1
python tools/one_kernel.py hessian 25216 3072 fp16 5 > $O/hessian_time.log 2>&1
python tools/one_kernel.py hessian 2056 6144 fp16 5 >> $O/hessian_time.log 2>&1
python tools/one_kernel.py hessian 25216 768 fp32 5 >> $O/hessian_time.log 2>&1
python tools/one_kernel.py obs 3072 768 fp16 3 > $O/obs_time.log 2>&1
python tools/one_kernel.py obs 6144 1408 fp16 3 >> $O/obs_time.log 2>&1
python tools/one_kernel.py obs 1408 6144 fp16 3 >> $O/obs_time.log 2>&1
ls -la $O
