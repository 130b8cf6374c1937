#!/bin/bash
mkdir -p gpurun_out
(echo "regs64"; python tools/rs_block.py; echo "regs80"; ECF_RS_REGS80=1 python tools/rs_block.py; echo "regs64"; python tools/rs_block.py; echo "regs80"; ECF_RS_REGS80=1 python tools/rs_block.py) 2>&1 | tee gpurun_out/rs_block_r4j.log
