#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "obs or sparsegpt or hessian or prepare_hinv" > gpurun_out/pytest_r2r.log 2>&1; tail -5 gpurun_out/pytest_r2r.log
python tools/sgpt_probe.py 2>&1 | tee gpurun_out/sgpt_probe.log
