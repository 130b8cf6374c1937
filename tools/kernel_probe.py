"""Quick device-side timing of the HBM-bound kernels at BLIP-2 / LLaMA shapes (development probe;
bench.py is the contract benchmark)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ecoflap_b200 import ops

dev = torch.device("cuda", 0)
PEAK = 6551.4
ONLY = set(sys.argv[1:])  # e.g. `kernel_probe.py row_select layer_thresh`; empty = everything


def want(k):
    return not ONLY or k in ONLY


def timeit(fn, reps=20, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = []
for (T, C, dt) in [] if not want('sqnorm') else [(2056, 1408, torch.float32), (2056, 6144, torch.float16), (512, 2048, torch.bfloat16), (32896, 1408, torch.float16), (32896, 1408, torch.float32), (32896, 6144, torch.float16), (8192, 2048, torch.bfloat16),
                   (8192, 5120, torch.bfloat16), (262144, 4096, torch.float16), (65536, 2048, torch.bfloat16)]:
    x = torch.randn(T, C, device=dev, dtype=dt)
    s = torch.zeros(C, device=dev)
    ms = timeit(lambda: ops.sqnorm_accum(x, s, 0.5, 0.5), flush=flush)
    gb = (T * C * x.element_size() + 8 * C) / 1e9
    out.append(dict(k="sqnorm", T=T, C=C, dt=str(dt), ms=ms, GBs=gb / ms * 1e3, frac=gb / ms * 1e3 / PEAK))
    print(out[-1], flush=True)
    del x
for (R, C, dt) in [] if not want('row_select') else [(2048, 2048, torch.bfloat16), (5120, 2048, torch.bfloat16), (2048, 5120, torch.bfloat16), (4096, 4096, torch.float16),
                   (11008, 4096, torch.float16), (4096, 11008, torch.float16), (3072, 768, torch.float16), (3072, 768, torch.float32), (16384, 2048, torch.bfloat16)]:
    W0 = (torch.randn(R, C, device=dev) * 0.02).to(dt)
    s = torch.rand(C, device=dev) + 0.1
    W = W0.clone()
    def run():
        ops.wanda_row_select_apply(W, s, C // 2)
    # note: after the first call half of W is zero; restore each rep outside the timed region
    ts = []
    for _ in range(12):
        W.copy_(W0); flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); ms = ts[len(ts) // 2]
    gb = (2 * R * C * W.element_size() + 4 * C) / 1e9
    out.append(dict(k="row_select", R=R, C=C, dt=str(dt), ms=ms, GBs=gb / ms * 1e3, frac=gb / ms * 1e3 / PEAK))
    print(out[-1], flush=True)
for (R, C, dt) in [] if not want('layer_thresh') else [(4224, 1408, torch.float16), (6144, 1408, torch.float16), (1408, 6144, torch.float16), (3072, 768, torch.float32)]:
    W0 = (torch.randn(R, C, device=dev) * 0.02).to(dt)
    s = torch.rand(C, device=dev) + 0.1
    W = W0.clone()
    ts = []
    for _ in range(12):
        W.copy_(W0); flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.wanda_layer_thresh_apply(W, s, R * C // 2); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); ms = ts[len(ts) // 2]
    gb = (2 * R * C * W.element_size() + 4 * C) / 1e9
    out.append(dict(k="layer_thresh", R=R, C=C, dt=str(dt), ms=ms, GBs=gb / ms * 1e3, frac=gb / ms * 1e3 / PEAK))
    print(out[-1], flush=True)
for (R, C, dt, n, m) in [] if not want('nm_select') else [(11008, 4096, torch.float16, 2, 4), (4096, 11008, torch.float16, 4, 8), (5120, 2048, torch.bfloat16, 2, 4)]:
    W0 = (torch.randn(R, C, device=dev) * 0.02).to(dt)
    s = torch.rand(C, device=dev) + 0.1
    W = W0.clone()
    ts = []
    for _ in range(12):
        W.copy_(W0); flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.wanda_nm_select_apply(W, s, n, m); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); ms = ts[len(ts) // 2]
    gb = (2 * R * C * W.element_size() + 4 * C) / 1e9
    out.append(dict(k="nm_select", R=R, C=C, dt=str(dt), nm=f"{n}:{m}", ms=ms, GBs=gb / ms * 1e3, frac=gb / ms * 1e3 / PEAK))
    print(out[-1], flush=True)
if not want('misc'):
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/kernel_probe_%s.json" % os.environ.get("PROBE_TAG", "x"), "w"), indent=1)
    sys.exit(0)
ts_ = [(torch.randn(4096, 4096, device=dev) * 0.02).to(torch.bfloat16) for _ in range(64)]
ms = timeit(lambda: ops.group_abs_reduce(ts_), flush=flush)
gb = sum(t.numel() * 2 for t in ts_) / 1e9
out.append(dict(k="group_reduce", ms=ms, GBs=gb / ms * 1e3, frac=gb / ms * 1e3 / PEAK)); print(out[-1])
w = (torch.randn(8192, 8192, device=dev)).to(torch.bfloat16); z = torch.randn_like(w)
ms = timeit(lambda: ops.zo_perturb(w, z, 1, 1e-3), flush=flush)
gb = w.numel() * 6 / 1e9
out.append(dict(k="zo_perturb", ms=ms, GBs=gb / ms * 1e3, frac=gb / ms * 1e3 / PEAK)); print(out[-1])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/kernel_probe.json", "w"), indent=1)
