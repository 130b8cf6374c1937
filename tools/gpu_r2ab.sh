#!/bin/bash
for st in 0 500 1500 3000; do
echo "== STAGGER=$st ns"
ECF_RS_STAGGER=$st python tools/rs_block.py
ECF_RS_STAGGER=$st python tools/kernel_probe.py row_select 2>&1 | grep "16384\|11008, 'C': 4096" | cut -c1-150
done
