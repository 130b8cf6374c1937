#!/bin/bash
# per-layer select: parity tests + phase timing of a ViT-g block
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "layer" > $O/pytest_lt.log 2>&1; tail -4 $O/pytest_lt.log
timeout 300 python tools/one_kernel.py layer_block 0 0 fp16 5 2>&1 | tail -4 | tee $O/lt_block.log
