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
if which == "layer_block":
    shapes = [(4224, 1408), (1408, 1408), (6144, 1408), (1408, 6144)]
    W0 = [(torch.randn(r, c, device=dev) * 0.02).half() for r, c in shapes]
    ss = [torch.rand(c, device=dev) + 0.1 for r, c in shapes]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(reps):
        Ws = [w.clone() for w in W0]
        if os.environ.get("FLUSH", "write") == "read":
            flush.sum()      # evicts with CLEAN lines (what a preceding streaming-read kernel leaves in L2)
        else:
            flush.zero_()    # evicts with DIRTY lines: the kernel's reads also pay their write-back
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.wanda_layer_thresh_apply_batched([(w, s, w.numel() // 2) for w, s in zip(Ws, ss)])
        e1.record()
        torch.cuda.synchronize()
        print("block select ms", e0.elapsed_time(e1), "phase us (sample, bracket, count, refine, apply):",
              [round(x, 1) for x in ops.layer_thresh_phase_times_us(dev)])
        print("   stamps us [start, P1+P2 end, =, P3 end, P4 end, P5 end, P3 loop, P3 walk, P3 flush, find0, walk1, flush1, find1, P1 sampled, P1 done]:",
              ops.layer_thresh_stamps_us(dev)[:15])
elif which in ("sqnorm_vit", "sqnorm_t5"):
    # the batched norm launch of one block forward at BLIP-2 shapes (R, C, dtype arguments are ignored)
    if which == "sqnorm_vit":
        shapes = [(2056, 1408, torch.float32), (2056, 1408, torch.float16), (2056, 1408, torch.float32), (2056, 6144, torch.float16)]
        share = [0, 1, 2, 3]
    else:
        shapes = [(512, 2048, torch.bfloat16), (512, 2048, torch.bfloat16), (512, 2048, torch.bfloat16), (512, 5120, torch.bfloat16)]
        share = [0, 0, 0, 1, 2, 2, 3]
    xs = [torch.randn(t, c, device=dev).to(d) for t, c, d in shapes]
    items = [(xs[i], torch.zeros(xs[i].shape[1], device=dev), 0.5, 0.5) for i in share]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(reps):
        flush.zero_()
        ops.sqnorm_accum_batched(items)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.sqnorm_accum_batched(items); ops.sqnorm_accum_batched(items); ops.sqnorm_accum_batched(items); ops.sqnorm_accum_batched(items); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 4)
    print(which, "us per launch (4 back to back, 1st cold):", sorted(ts)[len(ts) // 2] * 1e3)
elif which == "hessian":
    # R = tokens T, C = channels; X^T X on tcgen05 (A8)
    x = torch.randn(R, C, device=dev).to(dt)
    H = torch.zeros(C, C, device=dev)
    for _ in range(2):
        ops.hessian_accum(x, H, 0.5, 0.5)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.hessian_accum(x, H, 0.5, 0.5); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"hessian T={R} C={C} {sys.argv[4]}: {ms*1e3:.1f} us, {2*R*C*C/ms/1e9:.1f} TFLOP/s algorithmic (full product)")
elif which == "obs":
    # full OBS block loop (A10) on an R x C fp32 working copy with a well-conditioned upper factor
    from ecoflap_b200.accumulators import SparseGPT
    lin = torch.nn.Linear(C, R, bias=False).to(dev).to(dt)
    xs = torch.randn(max(2 * C, 4096), C, device=dev).to(dt)
    ts = []
    for _ in range(reps):
        with torch.no_grad():
            lin.weight.copy_(torch.randn(R, C, device=dev) * 0.02)
        sg = SparseGPT(lin)
        sg.add_batch(xs.unsqueeze(0))
        W = lin.weight.data.float().contiguous()
        Hinv, dead = sg.prepare_hinv(0.01)
        kth = [int(R * min(128, C - i) * 0.5) for i in range(0, C, 128)]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.obs_prune(W, Hinv, kth, 128); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"obs R={R} C={C}: {ms:.3f} ms, trailing-update {R*C*C/ms/1e9:.2f} TFLOP/s algorithmic")
elif which == "sqnorm":
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
            print("phase us (sample, bracket, count, refine, apply):", [round(x, 1) for x in ops.layer_thresh_phase_times_us(dev)])
torch.cuda.synchronize()
