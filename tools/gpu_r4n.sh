#!/bin/bash
mkdir -p gpurun_out
(for ns in 4 3.5 3 2.5; do echo "== ECF_LT_NSIGMA=$ns"; ECF_LT_NSIGMA=$ns timeout 300 python tools/lt_cut_probe.py vitg 2>&1 | grep -E "graph replay|fallback|first"; done) | tee gpurun_out/lt_nsigma_r4n.log
