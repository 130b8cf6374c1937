#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import numpy as np, torch
import ecoflap_oracle as orc
from ecoflap_b200 import ops
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
for (shapes, dt) in (([(70, 2048), (19, 5120)], torch.bfloat16), ([(40, 768), (33, 1024)], torch.float16), ([(21, 4096), (9, 3072)], torch.bfloat16)):
    Ws = [(torch.randn(r, c, generator=g) * 0.02).to(dt) for r, c in shapes]
    Ws[0][3] = 0; Ws[0][4, ::2] = 0; Ws[0][5] = Ws[0][5, 0]
    ss = [torch.rand(c, generator=g) + 0.1 for r, c in shapes]
    Wd = [w.clone().to(dev) for w in Ws]
    ops.wanda_row_select_apply_batched([(w, s.to(dev), w.shape[1] // 2) for w, s in zip(Wd, ss)])
    torch.cuda.synchronize()
    for w0, wd, s in zip(Ws, Wd, ss):
        want, _ = orc.wanda_prune_rows(w0.float().numpy(), s.numpy(), 0.5)
        assert np.array_equal(wd.float().cpu().numpy(), want)
print("sanitizer target ok")
PY
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/san.py > gpurun_out/memcheck_rs_r4o.log 2>&1
tail -4 gpurun_out/memcheck_rs_r4o.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python /tmp/san.py > gpurun_out/racecheck_rs_r4o.log 2>&1
tail -4 gpurun_out/racecheck_rs_r4o.log
