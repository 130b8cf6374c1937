"""Driver-side formats around the hot path: what ``LAVIS/evaluate_blip.py`` writes after
``pruner.prune()`` and reads back for evaluation, restated so that a run of this package leaves the same artefacts.

  * ``pruned_checkpoint/<job_id>.pth``   dense ``model.state_dict()`` with the zeros in place (evaluate_blip.py:438-448)
  * ``sparsity_dict/<job_id>.yaml``      the ``{parameter name: sparsity}`` dict of stage 1, only when it is a real dict
                                         -- the uniform module is not saved (:450-457)
  * ``training_statistics/<job_id>.yaml``  ``{"memory": peak GB, "time": seconds}`` (:459-472)
  * re-loading: the T5 part of a checkpoint by the ``t5_model.`` prefix (:344-351), the ViT part by ``visual.`` /
    ``visual_encoder.`` with unknown keys dropped (:353-389)
  * the kept-parameter percentage printed by the driver (:432-436), counted on the device (``ecf_count_zero``)

New here (no reference counterpart): a packed export of a pruned matrix -- bit mask + kept values -- for 2:4 / unstructured
sparse consumers.  Host-side torch code, not part of the hot path."""
from __future__ import annotations

import os
import time
from typing import Dict, Optional

import torch
import yaml

from . import ops


def kept_parameter_percentage(model) -> float:
    """``sum((p != 0).sum()) / sum(p.numel()) * 100`` over all parameters (evaluate_blip.py:340-342,432-436)."""
    total, zeros = 0, 0
    for p in model.parameters():
        total += p.numel()
        d = p.data if p.data.is_contiguous() else p.data.contiguous()
        zeros += int(ops.count_zero(d).item())
    return 100.0 * (total - zeros) / max(1, total)


def filter_eva_clip_checkpoint(state_dict):
    """EVA-CLIP runs save only the visual tower without its last block (evaluate_eva_clip.py:414-423): BLIP-2 uses the
    ViT-g's first 39 blocks, so ``blocks.39`` is dropped and everything outside ``visual.`` too."""
    return {k: v for k, v in state_dict.items() if "blocks.39" not in k and "visual." in k}


def save_pruning_outputs(model, job_id: str, sparsity_dict=None, start_time: Optional[float] = None, root: str = ".",
                         eva_clip: bool = False) -> Dict[str, str]:
    """Write the three artefacts of ``--save_pruned_model`` (evaluate_blip.py:438-472; ``eva_clip=True``: the filtered
    checkpoint of evaluate_eva_clip.py:410-428).  Returns their paths."""
    out = {}
    folder = os.path.join(root, "pruned_checkpoint")
    os.makedirs(folder, exist_ok=True)
    out["checkpoint"] = os.path.join(folder, job_id + ".pth")
    torch.save(filter_eva_clip_checkpoint(model.state_dict()) if eva_clip else model.state_dict(), out["checkpoint"])
    print(out["checkpoint"])
    if sparsity_dict is not None and isinstance(sparsity_dict, dict):
        folder = os.path.join(root, "sparsity_dict")
        os.makedirs(folder, exist_ok=True)
        out["sparsity_dict"] = os.path.join(folder, job_id + ".yaml")
        with open(out["sparsity_dict"], "w") as f:
            yaml.dump({k: float(v) for k, v in sparsity_dict.items()}, f)
    peak_memory = (torch.cuda.max_memory_allocated() / 1024 ** 2) / 1000 if torch.cuda.is_available() else 0.0
    processing_time = time.time() - start_time if start_time is not None else 0.0
    folder = os.path.join(root, "training_statistics")
    os.makedirs(folder, exist_ok=True)
    out["training_statistics"] = os.path.join(folder, job_id + ".yaml")
    with open(out["training_statistics"], "w") as f:
        yaml.dump({"memory": peak_memory, "time": processing_time}, f)
    return out


def load_sparsity_dict(path: str) -> Dict[str, float]:
    """``sparsity_dict`` constructor kwarg of the pruners: a yaml (or json, a yaml subset) file of name -> ratio."""
    with open(path) as f:
        return {k: float(v) for k, v in yaml.safe_load(f).items()}


def load_t5_pruned_checkpoint(model, path: str):
    """evaluate_blip.py:344-351: keep the ``t5_model.`` keys, strip the prefix, strict load into ``model.t5_model``."""
    sd = torch.load(path, map_location="cpu")
    sd = {k.replace("t5_model.", ""): v for k, v in sd.items() if k.startswith("t5_model")}
    model.t5_model.load_state_dict(sd)
    return model


def load_vit_pruned_checkpoint(model, path: str, interpolate_pos_embed=None):
    """evaluate_blip.py:353-389: the checkpoint's ViT keys (prefix ``visual.`` or ``visual_encoder.``) replace the matching
    entries of ``model.visual_encoder.state_dict()``; keys the encoder does not have are dropped; ``interpolate_pos_embed``
    (lavis.models.eva_vit) is applied when given."""
    sd = torch.load(path, map_location="cpu")
    prefix = None
    for cand in ("visual.", "visual_encoder."):
        if any(k.startswith(cand) for k in sd):
            prefix = cand
            break
    assert prefix is not None
    print(f"VIT checkpoint prefix: {prefix}")
    sd = {k.replace(prefix, ""): v for k, v in sd.items() if k.startswith(prefix)}
    full = model.visual_encoder.state_dict()
    for k, v in sd.items():
        if k in full:
            full[k] = v
    if interpolate_pos_embed is not None:
        interpolate_pos_embed(model.visual_encoder, full)
    model.visual_encoder.load_state_dict(full)
    return model


def pack_sparse(W: torch.Tensor):
    """Pruned [R, C] matrix -> (mask_bits uint8 [R, ceil(C/8)], values 1-D of the kept entries in row-major order).
    Bit j of byte v of a row = column 8v + j is kept (non-zero) -- the layout of the kernels' ``mask_bits``, inverted."""
    assert W.dim() == 2
    R, C = W.shape
    keep = W != 0
    pad = (-C) % 8
    kp = torch.nn.functional.pad(keep, (0, pad)) if pad else keep
    weights = (1 << torch.arange(8, device=W.device, dtype=torch.int32))
    bits = (kp.view(R, -1, 8).to(torch.int32) * weights).sum(-1).to(torch.uint8)
    return bits, W[keep]


def unpack_sparse(mask_bits: torch.Tensor, values: torch.Tensor, C: int) -> torch.Tensor:
    R = mask_bits.shape[0]
    shifts = torch.arange(8, device=mask_bits.device, dtype=torch.uint8)
    keep = ((mask_bits.unsqueeze(-1) >> shifts) & 1).bool().view(R, -1)[:, :C]
    W = torch.zeros(R, C, dtype=values.dtype, device=values.device)
    W[keep] = values
    return W
