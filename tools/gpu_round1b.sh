#!/bin/bash
# Round-1 (second pass) GPU evidence: parity tests, bench (both arms), ncu launch list of the bench command,
# ncu --set full of the norm kernel at the block-sweep shapes, select-kernel probes.
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 400 $O/bench.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -k regex:'sqnorm|row_select|layer_thresh' -c 2500 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/bench_under_ncu.log 2>&1
FULL="$NCU --set full --import-source on"
timeout 200 $FULL -k regex:sqnorm -s 2 -c 1 -f -o $O/sqnorm_vitg_block_r1b python tools/sq_probe.py ncu vitg_block > $O/ncu_sq1.log 2>&1
timeout 200 $FULL -k regex:sqnorm -s 2 -c 1 -f -o $O/sqnorm_t5enc_block_r1b python tools/sq_probe.py ncu t5_enc_block > $O/ncu_sq2.log 2>&1
timeout 200 python tools/sq_probe.py > $O/sq_probe.log 2>&1
timeout 200 python tools/rs_block.py > $O/rs_block.log 2>&1
ECF_RS_KEEP=1 timeout 200 python tools/rs_block.py >> $O/rs_block.log 2>&1
PROBE_TAG=r1b timeout 300 python tools/kernel_probe.py row_select layer_thresh sqnorm > $O/kernel_probe.log 2>&1
PROBE_TAG=r1b_keep1 ECF_RS_KEEP=1 timeout 300 python tools/kernel_probe.py row_select > $O/kernel_probe_keep1.log 2>&1
timeout 100 python tools/one_kernel.py layer_block 0 0 fp16 4 > $O/lt_block.log 2>&1
ls -la $O | head -40
