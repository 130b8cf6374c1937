#!/bin/bash
# sqnorm A/B sweep (tile width, occupancy, dedup) + its parity tests
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/sq_probe.log
for cfg in "" "ECF_SQ_TXL=5" "ECF_SQ_TXL=6" "ECF_SQ_OCC=3" "ECF_SQ_OCC=8" "ECF_SQ_DEDUP=0"; do
  echo "== cfg: $cfg" >> $O/sq_probe.log
  env $cfg timeout 300 python tools/sq_probe.py >> $O/sq_probe.log 2>&1
done
cat $O/sq_probe.log
