#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "hinv or sparsegpt or obs or hessian" > gpurun_out/pytest_r3k.log 2>&1
tail -3 gpurun_out/pytest_r3k.log
timeout 600 python tools/prune_wall.py sparsegpt > gpurun_out/prune_wall_sgpt_r3k.json 2> gpurun_out/prune_wall_sgpt_r3k.err
tail -c 600 gpurun_out/prune_wall_sgpt_r3k.json; tail -3 gpurun_out/prune_wall_sgpt_r3k.err
ECF_HINV_REFERENCE_ORDER=1 timeout 600 python tools/prune_wall.py sparsegpt > gpurun_out/prune_wall_sgpt_r3k_ref.json 2>/dev/null
tail -c 300 gpurun_out/prune_wall_sgpt_r3k_ref.json
