"""End-to-end golden fixtures: the UNMODIFIED reference pruner classes driving the tiny stand-in models on CPU.
Stores the pruned weight matrices, the (W_before, scaler_row, sparsity) triples the reference used at every
select step (captured by subclassing its WrappedGPT), and the sparsity dicts.  Run via tests/gen_golden.py."""
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


def _record_wrapped(module):
    """Replace module.WrappedGPT by a subclass that remembers every instance (to read scaler_row later)."""
    made = []
    base = module.WrappedGPT

    class Recording(base):
        def __init__(self, layer, *a, **k):
            super().__init__(layer, *a, **k)
            self._w_before = layer.weight.data.detach().clone()
            made.append(self)

    module.WrappedGPT = Recording
    return made, base


def _dump_steps(out, prefix, made, limit=6):
    for i, w in enumerate(made[:limit]):
        out[f"{prefix}__step{i}__W"] = w._w_before.float().numpy()
        out[f"{prefix}__step{i}__s"] = w.scaler_row.float().numpy()
        out[f"{prefix}__step{i}__Wafter"] = w.layer.weight.data.float().numpy()
    out[f"{prefix}__nsteps"] = np.int64(min(limit, len(made)))


def main():
    torch.set_num_threads(8)
    lavis = ref_loader.load_lavis_pruners()
    coop = ref_loader.load_coop_pruners()
    out = {}

    # 1. EVA-ViT, Wanda per-layer threshold, uniform 50 %
    made, base = _record_wrapped(lavis.wanda)
    m = cases.vit_model()
    p = lavis.wanda.VITLayerWandaPruner(model=m, data_loader=cases.vit_loader(), prune_spec="3-0.5-1.0-1.0",
                                        num_samples=16, model_prefix="visual")
    p.prune()
    lavis.wanda.WrappedGPT = base
    for k, v in cases.prunable_state(m).items():
        out[f"vit_wanda__{k}"] = v
    _dump_steps(out, "vit_wanda", made)

    # 2. T5, Wanda per-row, uniform 50 %
    made, base = _record_wrapped(lavis.wanda)
    m = cases.t5_model()
    p = lavis.wanda.T5LayerWandaPruner(model=m, data_loader=cases.t5_loader(), prune_spec="2-0.5-1.0-1.0",
                                       num_samples=16, model_prefix="t5_model")
    p.prune()
    lavis.wanda.WrappedGPT = base
    for k, v in cases.prunable_state(m).items():
        out[f"t5_wanda__{k}"] = v
    _dump_steps(out, "t5_wanda", made)

    # 3. BLIP-2, ECoFLaP zeroth-order + Wanda (stage 1 on the CPU generator -> ratios saved, stage 2 compared)
    np.random.seed(42)
    m = cases.blip2_model()
    p = lavis.wanda.BLIPT5LayerWandaPruner(
        model=m, data_loader=cases.blip2_loader(), t5_prune_spec="2-0.5-1.0-1.0", vit_prune_spec="3-0.5-1.0-1.0",
        t5_pruning_method="x", vit_pruning_method="x", num_samples=16, sparsity_ratio_granularity="block",
        max_sparsity_per_layer=0.6, score_method="MEZO-GradOnly_sum", num_data_first_stage=8, num_noise=1,
        noise_eps=1e-3)
    _, sd = p.prune()
    for k, v in cases.prunable_state(m).items():
        out[f"blip2_ecoflap__{k}"] = v
    out["blip2_ecoflap__sparsity_keys"] = np.array(list(sd.keys()))
    out["blip2_ecoflap__sparsity_vals"] = np.array([sd[k] for k in sd], dtype=np.float64)

    # 4. CLIP (CoOp), Wanda per-row 40 %, both towers
    m = cases.clip_model()
    p = coop.wanda.CLIPLayerWandaPruner(model=m, data_loader=cases.clip_loader(), language_prune_spec="1-0.6-1-1",
                                        visual_prune_spec="1-0.6-1-1", num_samples=16)
    from ecoflap_b200.synthetic import clip_forward_to_cache

    p.forward_to_cache = clip_forward_to_cache(cases.clip_class_tokens())
    p.prune()
    for k, v in cases.prunable_state(m).items():
        out[f"clip_wanda__{k}"] = v

    # 5. EVA-ViT, SparseGPT 40 % (batch size 1, as the reference asserts)
    m = cases.vit_model()
    p = lavis.sparsegpt.VITLayerSparseGPTPruner(model=m, data_loader=cases.vit_loader(batch=1, n=48),
                                                prune_spec="3-0.6-1.0-1.0", num_samples=48, model_prefix="visual")
    p.prune()
    for k, v in cases.prunable_state(m).items():
        out[f"vit_sparsegpt__{k}"] = v

    # 6. CLIP (CoOp), SparseGPT 40 %
    m = cases.clip_model()
    p = coop.sparsegpt.CLIPLayerSparseGPTPruner(model=m, data_loader=cases.clip_loader(n=96), language_prune_spec="1-0.6-1-1",
                                                visual_prune_spec="1-0.6-1-1", num_samples=96)
    p.forward_to_cache = clip_forward_to_cache(cases.clip_class_tokens())
    p.prune()
    for k, v in cases.prunable_state(m).items():
        out[f"clip_sparsegpt__{k}"] = v

    path = os.path.join(GOLD, "e2e_pruners.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
