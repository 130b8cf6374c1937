"""Stage-1 fixtures (A12 zeroth-order, A13 first-order) from the UNMODIFIED reference LayerSparsity on CPU, on the
toy BLIP-2 stand-in of tests/e2e_cases.py.  Run in the build container only:  python tests/gen_golden_stage1.py
Stores, per score method: the per-layer score sums (fp64 sums of the reference's per-element score tensors; for
GradOnly the g-hat itself) and the sparsity dict return_sparsity() produced from them."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)
import e2e_cases as cases  # noqa: E402
import ref_loader  # noqa: E402

GOLD = os.path.join(HERE, "golden")
METHODS = ("GradMagAbs_sum", "GradMagSquare_avg", "GradOnly_sum", "MEZO-GradOnly_sum", "MEZO-GradMagAbs_sum")


def main():
    torch.set_num_threads(8)
    lavis = ref_loader.load_lavis_pruners()
    out = {}
    for method in METHODS:
        np.random.seed(42)
        torch.manual_seed(0)
        m = cases.blip2_model()
        p = lavis.wanda.BLIPT5LayerWandaPruner(
            model=m, data_loader=cases.blip2_loader(), t5_prune_spec="2-0.5-1.0-1.0", vit_prune_spec="3-0.5-1.0-1.0",
            t5_pruning_method="x", vit_pruning_method="x", num_samples=16, sparsity_ratio_granularity="block",
            max_sparsity_per_layer=0.6, score_method=method, num_data_first_stage=8, num_noise=1, noise_eps=1e-3)
        rec = {}
        cls = lavis.LayerSparsity
        saved = {fn: getattr(cls, fn) for fn in ("compute_importance_scores", "compute_importance_scores_mezo")}
        for fn, orig in saved.items():

            def wrap(self, mapping, _orig=orig):
                res = _orig(self, mapping)
                rec["scores"] = res
                return res

            setattr(cls, fn, wrap)
        try:
            for prm in m.parameters():
                prm.requires_grad = True
            sd = p.get_sparsity(0.5, sparsity_ratio_granularity="block")
        finally:
            for fn, orig in saved.items():
                setattr(cls, fn, orig)
        keys = list(sd.keys())
        out[f"{method}__keys"] = np.array(keys)
        out[f"{method}__sums"] = np.array([float(rec["scores"][k].double().sum()) for k in keys], dtype=np.float64)
        out[f"{method}__res"] = np.array([sd[k] for k in keys], dtype=np.float64)
        print(method, len(keys), out[f"{method}__sums"][:3], out[f"{method}__res"][:3])
    out["cases"] = np.array(METHODS)
    path = os.path.join(GOLD, "stage1_scores.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
