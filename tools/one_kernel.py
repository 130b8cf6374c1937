"""Launch one hot-path kernel a few times at a named shape (target for `ncu --set full -k regex:...`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ecoflap_b200 import ops

dev = torch.device("cuda", 0)
which = sys.argv[1]
R, C = int(sys.argv[2]), int(sys.argv[3])
dt = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[sys.argv[4]]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
torch.manual_seed(0)
if which == "sqnorm":
    x = torch.randn(R, C, device=dev).to(dt)
    s = torch.zeros(C, device=dev)
    for _ in range(reps):
        ops.sqnorm_accum(x, s, 0.5, 0.5)
else:
    W0 = (torch.randn(R, C, device=dev) * 0.02).to(dt)
    s = torch.rand(C, device=dev) + 0.1
    for _ in range(reps):
        W = W0.clone()
        if which == "row_select":
            ops.wanda_row_select_apply(W, s, C // 2)
        else:
            ops.wanda_layer_thresh_apply(W, s, R * C // 2)
torch.cuda.synchronize()
