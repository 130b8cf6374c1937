#!/bin/bash
# Round-1 final GPU evidence: parity tests, smoke, bench (both arms), ncu launch list of the bench command, kernel probes.
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -k regex:'sqnorm|row_select|layer_thresh' -c 2500 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/bench_under_ncu.log 2>&1
PROBE_TAG=r1c timeout 300 python tools/kernel_probe.py > $O/kernel_probe.log 2>&1
timeout 200 python tools/rs_block.py > $O/rs_block.log 2>&1
ls $O | head -50
