"""Generates tests/golden/*.npz by running the UNMODIFIED reference classes (imported from
/root/reference through oracle/ref_loader.py) on seeded synthetic inputs.

Run in the build container only:   python tests/gen_golden.py
The fixtures are committed; the GPU box never needs the reference tree.

Each fixture stores the exact inputs (low-precision tensors as float32 + a dtype tag) and the
outputs the reference produced, so that both the numpy oracle (CPU tests) and the CUDA path
(GPU tests) can be checked against the reference itself.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
sys.path.insert(0, os.path.join(HERE, ".."))
import ref_loader  # noqa: E402

GOLD = os.path.join(HERE, "golden")
TD = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}


def f32(t):
    return t.detach().float().cpu().numpy()


def save(name, **arrays):
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


# ---------------------------------------------------------------- A1 norm accumulator
def gen_norm(ref):
    out = {}
    g = torch.Generator().manual_seed(1)
    cases = [
        ("fp32_3d", "fp32", [(4, 7, 48), (4, 7, 48), (3, 7, 48)]),  # ragged last batch
        ("fp16_3d", "fp16", [(2, 33, 64), (2, 33, 64)]),
        ("bf16_3d", "bf16", [(8, 5, 40), (8, 5, 40), (8, 5, 40)]),
        ("fp16_2d", "fp16", [(19, 24), (19, 24)]),  # 2-D input is treated as B = 1
        ("fp32_outlier", "fp32", [(2, 16, 32), (2, 16, 32)]),
    ]
    names = []
    for name, dt, shapes in cases:
        layer = nn.Linear(shapes[0][-1], 8)
        acc = ref.WrappedGPT(layer)
        for i, shp in enumerate(shapes):
            x = torch.randn(*shp, generator=g)
            if name == "fp32_outlier":
                x[..., 3] *= 30.0
                x[..., 5] = 0.0
            x = x.to(TD[dt])
            acc.add_batch(x, None)
            out[f"{name}__x{i}"] = f32(x)
            out[f"{name}__s{i}"] = acc.scaler_row.clone().numpy()
            out[f"{name}__n{i}"] = np.int64(acc.nsamples)
        out[f"{name}__dtype"] = np.array(dt)
        out[f"{name}__nb"] = np.int64(len(shapes))
        names.append(name)
    out["cases"] = np.array(names)
    save("norm_accum", **out)


# ---------------------------------------------------------------- A8 Hessian accumulator
def gen_hessian(ref):
    out = {}
    g = torch.Generator().manual_seed(2)
    cases = [
        ("fp32_b1", "fp32", [(1, 9, 32)] * 3),
        ("fp16_b4", "fp16", [(4, 17, 64), (4, 17, 64), (2, 17, 64)]),
        ("bf16_2d", "bf16", [(21, 48), (21, 48)]),
    ]
    names = []
    for name, dt, shapes in cases:
        layer = nn.Linear(shapes[0][-1], 8)
        acc = ref.SparseGPT(layer)
        for i, shp in enumerate(shapes):
            x = torch.randn(*shp, generator=g).to(TD[dt])
            acc.add_batch(x, None)
            out[f"{name}__x{i}"] = f32(x)
            out[f"{name}__H{i}"] = acc.H.clone().numpy()
        out[f"{name}__dtype"] = np.array(dt)
        out[f"{name}__nb"] = np.int64(len(shapes))
        names.append(name)
    out["cases"] = np.array(names)
    save("hessian_accum", **out)


# ---------------------------------------------------------------- A9/A10 fasterprune
def gen_obs(ref):
    out = {}
    g = torch.Generator().manual_seed(3)
    cases = [
        ("fp32_64x256_s40", "fp32", 64, 256, 0.4, False, 0, 0),
        ("fp16_48x320_s50", "fp16", 48, 320, 0.5, False, 0, 0),  # ragged last block (320 = 2*128+64)
        ("fp32_32x128_dead", "fp32", 32, 128, 0.5, True, 0, 0),  # dead column path
        ("fp32_40x256_2of4", "fp32", 40, 256, 0.5, False, 2, 4),  # n:m branch (sparsegpt_pruner.py:195-198)
        ("fp16_24x384_4of8", "fp16", 24, 384, 0.5, False, 4, 8),
    ]
    names = []
    for name, dt, R, C, s, dead, pn, pm in cases:
        layer = nn.Linear(C, R, bias=False)
        with torch.no_grad():
            layer.weight.copy_(torch.randn(R, C, generator=g) * 0.02)
        layer = layer.to(TD[dt])
        acc = ref.SparseGPT(layer)
        for i in range(4):
            x = torch.randn(1, 3 * C // 2, C, generator=g)
            x[..., 1] *= 8.0
            if dead:
                x[..., 7] = 0.0
            x = x.to(TD[dt])
            acc.add_batch(x, None)
        out[f"{name}__H"] = acc.H.clone().numpy()
        out[f"{name}__W"] = f32(layer.weight.data)
        acc.fasterprune(s, prune_n=pn, prune_m=pm, percdamp=0.01, blocksize=128)
        out[f"{name}__Wout"] = f32(layer.weight.data)
        out[f"{name}__nm"] = np.array([pn, pm], dtype=np.int64)
        out[f"{name}__dtype"] = np.array(dt)
        out[f"{name}__s"] = np.float64(s)
        names.append(name)
    out["cases"] = np.array(names)
    save("obs_prune", **out)


# ---------------------------------------------------------------- A11 zeroth-order perturbation
def gen_zo(ref):
    out = {}
    names = []
    ls = ref.LayerSparsity.__new__(ref.LayerSparsity)
    for dt in ("fp32", "fp16", "bf16"):
        torch.manual_seed(11)
        p = nn.Parameter((torch.randn(37, 53) * 0.02).to(TD[dt]), requires_grad=False)
        out[f"{dt}__W0"] = f32(p.data)
        seed = 123456
        # z exactly as the reference draws it (CPU generator)
        torch.manual_seed(seed)
        z = torch.normal(mean=0, std=1, size=p.data.size(), dtype=p.data.dtype)
        out[f"{dt}__z"] = f32(z)
        for step, sc in enumerate((1, -2, 1)):
            ls.zo_perturb_parameters([p], random_seed=seed, scaling_factor=sc, zo_eps=1e-3)
            out[f"{dt}__W{step + 1}"] = f32(p.data)
        names.append(dt)
    out["cases"] = np.array(names)
    out["eps"] = np.float64(1e-3)
    save("zo_perturb", **out)


# ---------------------------------------------------------------- A15 allocator
def gen_alloc(ref):
    ls = ref.LayerSparsity.__new__(ref.LayerSparsity)
    out = {}
    names = []
    rng = np.random.default_rng(5)

    def run(name, scores, sizes, keep, maxsp):
        gs = {f"g{i}": torch.tensor(float(s), dtype=torch.float32) for i, s in enumerate(scores)}
        gn = {f"g{i}": int(n) for i, n in enumerate(sizes)}
        res = ls.compute_the_sparsity_per_group(int(keep), gs, gn, max_sparsity_per_layer=maxsp)
        out[f"{name}__scores"] = np.array([float(v) for v in gs.values()], dtype=np.float32)
        out[f"{name}__sizes"] = np.array(list(gn.values()), dtype=np.int64)
        out[f"{name}__keep"] = np.int64(keep)
        out[f"{name}__maxsp"] = np.float64(maxsp)
        out[f"{name}__res"] = np.array([res[k] for k in gn], dtype=np.float64)
        names.append(name)

    # the SURVEY toy KAT: scores {1,3}, sizes {100,200}, keep 150, max 0.6
    run("toy", [1.0, 3.0], [100, 200], 150, 0.6)
    run("toy_equal", [1.0, 1.0, 1.0], [10, 20, 30], 30, 0.8)
    # BLIP-2 scale: 39 ViT-g blocks + 24 T5 enc + 24 T5 dec
    vit = 4224 * 1408 + 1408 * 1408 + 2 * 6144 * 1408
    enc = 4 * 2048 * 2048 + 3 * 5120 * 2048
    dec = 8 * 2048 * 2048 + 3 * 5120 * 2048
    sizes = [vit] * 39 + [enc] * 24 + [dec] * 24
    for t in range(4):
        scores = np.abs(rng.normal(1.0, 0.6, size=len(sizes))).astype(np.float32) + 1e-3
        keep = int(sum(sizes) * (1 - 0.5))
        run(f"blip2_{t}", scores, sizes, keep, 0.6)
    # 'avg' style tiny scores, saturating groups, many small groups
    for t in range(6):
        G = int(rng.integers(3, 40))
        sizes = rng.integers(1000, 3_000_000, size=G)
        scores = (rng.random(G).astype(np.float32) ** 3) * (10.0 ** rng.integers(-6, 3))
        sp = float(rng.choice([0.3, 0.4, 0.5, 0.6]))
        maxsp = float(min(0.95, sp + rng.choice([0.0, 0.1, 0.2, 0.3])))
        keep = int(int(sizes.sum()) * (1 - sp))
        run(f"rand_{t}", scores, sizes, keep, maxsp)
    out["cases"] = np.array(names)
    save("allocator", **out)


# ---------------------------------------------------------------- A14 + return_sparsity
def gen_return_sparsity(ref):
    """LayerSparsity.return_sparsity with a pre-filled importance_measure (skips the loss loop)."""
    out = {}
    names = []
    torch.manual_seed(7)

    class Toy(nn.Module):
        def __init__(self):
            super().__init__()
            self.blocks = nn.ModuleList(
                [nn.ModuleDict({"a": nn.Linear(24, 40, bias=False), "b": nn.Linear(40, 24, bias=False)}) for _ in range(5)]
            )

    for method in ("MEZO-GradOnly_sum", "MEZO-GradMagAbs_sum", "MEZO-GradMagSquare_avg", "MEZO-GradOnly_avg"):
        m = Toy()
        mapping = {k: ".".join(k.split(".")[:2]) for k, _ in m.named_parameters()}
        ls = ref.LayerSparsity(m, None, None, 8, 0.5, 0.7, method, 1, 1e-3, mapping)
        ghat = {k: float(abs(torch.randn(()))) + 0.05 for k in mapping}
        comp = method.split("_")[0]
        imp = {}
        for k, v in m.named_parameters():
            g = torch.FloatTensor([ghat[k]]).abs()
            if comp == "MEZO-GradOnly":
                imp[k] = g.abs()
            elif comp == "MEZO-GradMagAbs":
                imp[k] = v.cpu().data.float().abs() * g.abs()
            else:
                imp[k] = v.cpu().data.float() ** 2 * g ** 2
        ls.importance_measure = imp
        res = ls.return_sparsity()
        keys = list(mapping)
        out[f"{method}__keys"] = np.array(keys)
        out[f"{method}__groups"] = np.array([mapping[k] for k in keys])
        out[f"{method}__ghat"] = np.array([ghat[k] for k in keys], dtype=np.float64)
        out[f"{method}__numel"] = np.array([dict(m.named_parameters())[k].numel() for k in keys], dtype=np.int64)
        out[f"{method}__impsum"] = np.array([float(imp[k].sum()) for k in keys], dtype=np.float32)
        for k in keys:
            out[f"{method}__W__{k}"] = f32(dict(m.named_parameters())[k])
        out[f"{method}__res"] = np.array([res[k] for k in keys], dtype=np.float64)
        names.append(method)
    out["cases"] = np.array(names)
    out["sparsity"] = np.float64(0.5)
    out["maxsp"] = np.float64(0.7)
    save("return_sparsity", **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    ref = ref_loader.load_lavis_pruners()
    gen_norm(ref)
    gen_hessian(ref)
    gen_obs(ref)
    gen_zo(ref)
    gen_alloc(ref)
    gen_return_sparsity(ref)
    try:
        import gen_golden_e2e
    except ImportError:
        return

    gen_golden_e2e.main()


if __name__ == "__main__":
    main()
