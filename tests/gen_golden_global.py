"""Generates tests/golden/global_mask.npz from the UNMODIFIED reference (LAVIS global_pruner.py:116-157), loaded through
oracle/ref_loader.py.  Run in the build container only (needs /root/reference):  python tests/gen_golden_global.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import ref_loader  # noqa: E402


def main():
    ns = ref_loader.load_lavis_pruners()
    G = ns.glob.BLIPT5GlobalPruner
    rng = np.random.default_rng(11)
    out = {}
    shapes = {"visual_encoder.blocks.0.attn.qkv.weight": (24, 16), "t5_model.encoder.block.0.layer.0.SelfAttention.q.weight": (16, 16),
              "t5_model.decoder.block.1.layer.2.DenseReluDense.wo.weight": (16, 40)}
    for case, (p, max_sp, ties) in enumerate([(0.5, 0.8, False), (0.7, 0.6, False), (0.3, 1.0, False), (0.5, 0.8, True)]):
        scores = {}
        for i, (k, shp) in enumerate(shapes.items()):
            v = np.abs(rng.standard_normal(shp)).astype(np.float32) * (1.0 + i)
            if ties:
                v[:, :4] = v[0, 0]
                v[1] = 0.0
            scores[k] = v
        ref = G.get_mask(None, {k: torch.from_numpy(v.copy()) for k, v in scores.items()}, p, max_sp)
        lw = G.get_layerwise_mask(None, {k: torch.from_numpy(v.copy()) for k, v in scores.items()}, p)
        out[f"c{case}__p"] = np.float64(p)
        out[f"c{case}__max_sp"] = np.float64(max_sp)
        for i, k in enumerate(shapes):
            out[f"c{case}__score{i}"] = scores[k]
            out[f"c{case}__mask{i}"] = ref[k].numpy()
            out[f"c{case}__lw{i}"] = lw[k].numpy()
    out["names"] = np.array(list(shapes))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "global_mask.npz"), **out)
    print("wrote tests/golden/global_mask.npz with", len(out), "arrays")
    e2e(ns)


def e2e(ns):
    """The unmodified reference classes end to end on the toy BLIP-2 (CPU, fp32): pruned weights of the global pruners
    (magnitude: global x 3 iterations, model-level global, layer-wise; first-order |w||g|: global x 2 iterations) and the
    sparsity dict of the 'RealGradMagAbs_sum' ratio oracle."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import e2e_cases as cases

    out = {}
    spec = dict(t5_prune_spec="2-0.5-1.0-1.0", vit_prune_spec="3-0.5-1.0-1.0", t5_pruning_method="x", vit_pruning_method="x")
    runs = [("mag_global3", ns.glob.BLIPT5GlobalMagPruner, dict(is_global=True, iteration=3)),
            ("mag_permodel", ns.glob.BLIPT5GlobalMagPruner, dict(is_global=True, prune_per_model=True, iteration=1)),
            ("mag_layerwise", ns.glob.BLIPT5GlobalMagPruner, dict(is_global=False, iteration=2)),
            ("gradmagabs_global2", ns.glob.BLIPT5GlobalGradMagAbsPruner, dict(is_global=True, iteration=2, num_samples=8))]
    for tag, cls, kw in runs:
        torch.manual_seed(0)
        m = cases.blip2_model()
        p = cls(model=m, data_loader=cases.blip2_loader(), **spec, **kw)
        p.prune()
        for k, v in cases.prunable_state(m).items():
            out[f"{tag}__{k}"] = np.packbits(v == 0)  # the zero pattern (weights = original * mask, the model is reproducible)
        print(tag, {k: float((v == 0).mean()) for k, v in list(cases.prunable_state(m).items())[:3]})
    torch.manual_seed(0)
    m = cases.blip2_model()
    p = ns.wanda.BLIPT5LayerWandaPruner(model=m, data_loader=cases.blip2_loader(), num_samples=16,
                                        sparsity_ratio_granularity="block", max_sparsity_per_layer=0.6,
                                        score_method="RealGradMagAbs_sum", num_data_first_stage=8, **spec)
    for prm in m.parameters():
        prm.requires_grad = True
    sd = p.get_sparsity(0.5, sparsity_ratio_granularity="block")
    out["real__keys"] = np.array(list(sd.keys()))
    out["real__vals"] = np.array([sd[k] for k in sd], dtype=np.float64)
    print("real", len(sd), list(sd.items())[:2])
    path = os.path.join(ROOT, "tests", "golden", "global_e2e.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
