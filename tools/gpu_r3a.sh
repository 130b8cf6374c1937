#!/bin/bash
mkdir -p gpurun_out
./tools/micro/cmp_micro.bin > gpurun_out/cmp_micro.log 2>&1
cat gpurun_out/cmp_micro.log
