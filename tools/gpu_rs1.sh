#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/rs_sweep.log
for cfg in "" "ECF_RS_NVMAX=4" "ECF_RS_NVMAX=2" "ECF_RS_KEEP=1" "ECF_RS_NVMAX=4 ECF_RS_KEEP=1"; do
  echo "== cfg: $cfg" >> $O/rs_sweep.log
  env $cfg timeout 300 python tools/rs_block.py >> $O/rs_sweep.log 2>&1
done
cat $O/rs_sweep.log
