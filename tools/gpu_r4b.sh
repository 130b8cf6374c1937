#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/prof_sweep.py sparsegpt 2>/dev/null | grep -v "^$" > gpurun_out/prof_sweep_r4b.log 2>&1
cut -c1-170 gpurun_out/prof_sweep_r4b.log | head -120
