#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/rs_sweep.log
for cfg in "ECF_RS_PREFETCH=0" "ECF_RS_PREFETCH=1" "ECF_RS_PREFETCH=1 ECF_RS_KEEP=1"; do
  echo "== cfg: $cfg" >> $O/rs_sweep.log
  env $cfg timeout 300 python tools/rs_block.py >> $O/rs_sweep.log 2>&1
  env $cfg PROBE_TAG=x timeout 300 python tools/kernel_probe.py row_select 2>&1 | cut -c1-150 | grep "11008\|16384\|'R': 2048, 'C': 2048" >> $O/rs_sweep.log
done
cat $O/rs_sweep.log
