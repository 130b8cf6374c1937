#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r2o.log 2>&1
grep -v "sparsity:" gpurun_out/pytest_gpu_r2o.log | tail -15
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
