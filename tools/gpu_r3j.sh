#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/chol_probe.py 2>&1 | tee gpurun_out/chol_probe.log
