#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "row_select" > gpurun_out/pytest_r4r.log 2>&1
tail -1 gpurun_out/pytest_r4r.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python tools/rs_block.py
