"""UPop entry point fixture: the UNMODIFIED reference BLIPBertLayerWandaPruner (UPop/pruners/wanda_pruner.py:600-834,
task="coco", tuple batches) on the toy BLIP stand-in, CPU.  Requested granularity "block" -- which the reference's
positional-argument quirk (:707-717) turns into uniform sparsity.  Run in the build container only."""
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


def main():
    torch.set_num_threads(8)
    upop = ref_loader.load_upop_pruners()
    out = {}
    for gran in (None, "block"):
        m = cases.caption_model()
        p = upop.wanda.BLIPBertLayerWandaPruner(
            model=m, data_loader=cases.caption_loader(), bert_prune_spec="2-0.5-1.0-1.0", vit_prune_spec="2-0.5-1.0-1.0",
            bert_model_prefix="text_decoder", vit_model_prefix="visual_encoder", num_samples=16, task="coco",
            sparsity_ratio_granularity=gran, max_sparsity_per_layer=0.6, score_method="GradMagAbs_sum",
            num_data_first_stage=8)
        _, sd = p.prune()
        tag = "uniform" if gran is None else "block"
        for k, v in cases.prunable_state(m).items():
            out[f"upop_{tag}__{k}"] = v
        out[f"upop_{tag}__sd_is_uniform"] = np.bool_(sd is None or not isinstance(sd, dict))
        print(tag, type(sd), {k: float((v == 0).mean()) for k, v in list(cases.prunable_state(m).items())[:4]})
    path = os.path.join(HERE, "golden", "e2e_upop.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
