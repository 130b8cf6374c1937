#!/bin/bash
mkdir -p gpurun_out
(echo "ECF_RS_TMA=0"; ECF_RS_TMA=0 timeout 300 python tools/rs_block.py
echo "ECF_RS_CORUN=0"; ECF_RS_CORUN=0 timeout 300 python tools/rs_block.py
for sh in 2 3; do echo "CORUN short=$sh"; ECF_RS_CORUN_SHORT=$sh timeout 300 python tools/rs_block.py; done
echo "enc only corun0"; RS_BLOCKS=enc,enc,enc,enc ECF_RS_CORUN=0 timeout 300 python tools/rs_block.py
echo "enc only corun1"; RS_BLOCKS=enc,enc,enc,enc timeout 300 python tools/rs_block.py
) 2>&1 | tee gpurun_out/rs_block_r3o.log
