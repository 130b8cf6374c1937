#!/bin/bash
# per-layer select v2: parity tests + phase timing of a ViT-g block
mkdir -p gpurun_out; O=gpurun_out

timeout 300 python tools/one_kernel.py layer_block 0 0 fp16 6 2>&1 | tee $O/lt_block.log

