"""Per-layer select probe: exactness against torch.kthvalue on the GPU and graph-replay timing of one block
(ViT-g: qkv, proj, fc1, fc2) after an L2 flush -- the way bench.py times roofline.kernels.layer_thresh.
usage: python tools/lt_cut_probe.py [vitg|llama|t5] [fp16|bf16] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ecoflap_b200 import ops

dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "vitg"
dt = {"fp16": torch.float16, "bf16": torch.bfloat16}[sys.argv[2] if len(sys.argv) > 2 else "fp16"]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
shapes = {"vitg": [(4224, 1408), (1408, 1408), (6144, 1408), (1408, 6144)],
          "llama": [(11008, 4096)], "t5": [(2048, 2048), (5120, 2048), (2048, 5120)]}[which]
torch.manual_seed(0)
W0 = [(torch.randn(r, c, device=dev) * 0.02).to(dt) for r, c in shapes]
ss = []
for r, c in shapes:
    s = torch.rand(c, device=dev) + 0.1
    s[::97] *= 900.0   # outlier channels
    s[5] = 0.0         # dead channel
    ss.append(s)
sp = [0.5, 0.41999998688697815, 0.6, 0.3][: len(shapes)]
idx = [int(w.numel() * s) for w, s in zip(W0, sp)]
Ws = [w.clone() for w in W0]
th = [torch.zeros(1, device=dev) for _ in shapes]
items = [(w, s, k, t) for w, s, k, t in zip(Ws, ss, idx, th)]
ops.wanda_layer_thresh_apply_batched(items)
print("fallback after first call:", ops.layer_thresh_last_fallback(dev))
ok = True
for i, (r, c) in enumerate(shapes):
    score = W0[i].float().abs() * ss[i].sqrt()[None, :]
    kth = torch.kthvalue(score.flatten(), idx[i] + 1).values
    mask = score <= kth
    e1 = float(th[i].item()) == float(kth.item())
    e2 = torch.equal(Ws[i], torch.where(mask, torch.zeros_like(W0[i]), W0[i]))
    print(f"  {r}x{c}: thres {'ok' if e1 else 'WRONG'} ({th[i].item():.9g} vs {kth.item():.9g}), weights {'ok' if e2 else 'WRONG'}, pruned {int(mask.sum())} of {mask.numel()}")
    ok = ok and e1 and e2
print("EXACT" if ok else "MISMATCH")

scratch = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    ops.wanda_layer_thresh_apply_batched(items)
ts = []
for _ in range(reps):
    for w, w0 in zip(Ws, W0):
        w.copy_(w0)
    scratch.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
nbytes = sum(2 * w.numel() * w.element_size() + 4 * w.shape[1] for w in W0)
best = min(ts)
print(f"{which} {sys.argv[2] if len(sys.argv) > 2 else 'fp16'}: graph replay us {[round(t, 1) for t in ts]}  best {best:.1f} us  "
      f"{nbytes / best / 1e6:.2f} TB/s algorithmic  fallback={ops.layer_thresh_last_fallback(dev)}")
# warm-L2 variant (weights restored right before: what the sweep sees after a block forward)
ts = []
for _ in range(reps):
    for w, w0 in zip(Ws, W0):
        w.copy_(w0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
print(f"   warm L2: best {min(ts):.1f} us")

# clean flush: the write flush leaves L2 full of DIRTY lines whose write-back competes with the kernel's own traffic; a
# streaming read after it leaves CLEAN lines (what a preceding norm kernel leaves behind in the real sweep)
scratch2 = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(reps):
    for w, w0 in zip(Ws, W0):
        w.copy_(w0)
    scratch.zero_()
    scratch2.sum()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
print(f"   clean flush (write + read): best {min(ts):.1f} us")
ge = torch.cuda.CUDAGraph()
tiny = torch.zeros(1, device=dev)
with torch.cuda.graph(ge):
    tiny.add_(1)
ts = []
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ge.replay(); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
print(f"   (one trivial kernel in a graph, same timing: {min(ts):.1f} us)")

# cumulative chain timing: K0 | K0+K1 | K0+K1+K3 | all four (ECF_LT_STOP is read per call, i.e. at capture time)
for stop in (1, 2, 3, 4):
    os.environ["ECF_LT_STOP"] = str(stop)
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        ops.wanda_layer_thresh_apply_batched(items)
    ts = []
    for _ in range(reps):
        for w, w0 in zip(Ws, W0):
            w.copy_(w0)
        scratch.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g2.replay(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    print(f"   first {stop} kernel(s): best {min(ts):.1f} us")
os.environ.pop("ECF_LT_STOP")
