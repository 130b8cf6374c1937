#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/prof_sweep.py ecoflap 2>/dev/null | grep -v "^$" > gpurun_out/prof_eco_r4e.log 2>&1
cut -c1-170 gpurun_out/prof_eco_r4e.log | head -75
