#!/bin/bash
for keep in 1 0; do
echo "== KEEP=$keep (4 CTAs/SM bound for KEEP=0)"
ECF_RS_KEEP=$keep python tools/rs_block.py
ECF_RS_KEEP=$keep python tools/kernel_probe.py row_select 2>&1 | grep row_select | cut -c1-160
done
