"""Generates tests/golden/global_mask.npz from the UNMODIFIED reference (LAVIS global_pruner.py:116-157), loaded through
oracle/ref_loader.py.  Run in the build container only (needs /root/reference):  python tests/gen_golden_global.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
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


if __name__ == "__main__":
    main()
