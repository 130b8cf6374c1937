"""Global-pruner baselines: ``blipt5_global_mag_pruner``, ``blipt5_global_gradmagabs_pruner`` and
``blipt5_global_mezo_pruner`` with the reference's constructor and ``prune()`` contract
(LAVIS/lavis/compression/pruners/global_pruner.py:56-388), and the mask step the ``Real*`` score methods of
``LayerSparsity`` reuse as a ratio oracle (layer_single_base_pruner.py:156-245).

One threshold over the scores of ALL prunable tensors (or of a tower, or one per tensor), ``iteration`` rounds with target
``p_i = p ** (iteration / i)``, weights multiplied by the 0/1 mask after each round.  The reference builds every score
tensor on the CPU in fp32 and runs ``torch.topk`` over their concatenation (3.7 G elements for BLIP-2); here the scores
are recomputed from the weights (and the accumulated |grad| sums, kept on the device) inside a radix select over a device
pointer table -- ``ops.GlobalTable`` / csrc/global_select.cu.  Reference behaviour that is kept on purpose:
  * the magnitude score is the SIGNED weight (global_pruner.py:249-251 has no abs), so the most negative weights go first;
  * ``prune()`` calls the mask step with max_sparsity_per_layer = 1.0, i.e. without the per-layer protection
    (global_pruner.py:239-242); the protection itself (the int(numel * (1 - max)) largest scores of a tensor count as
    finfo.max) is implemented for callers that pass another value;
  * a target that rounds to zero elements raises IndexError (``threshold[-1]`` of an empty top-k);
  * the zeroth-order variant has ONE score per tensor, so its mask switches whole tensors off.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from ..registry import registry
from .base import LayerWiseBasePruner, print_time
from .losses import loss_vision_language


def accumulate_abs_grads(model, data_loader, loss_func, names, params, num_samples, square=False):
    """sum over the first-stage batches of |dL/dW| (or (dL/dW)^2) per element, fp32, ON THE DEVICE, and the number of
    batches (global_pruner.py:262-296, layer_single_base_pruner.py:430-462; the reference adds them up on the CPU)."""
    device = next(iter(model.parameters())).device
    G = [torch.zeros(p.shape, dtype=torch.float32, device=p.device) for p in params]
    seen, nb = 0, 0
    for d in data_loader:
        if seen >= num_samples:
            break
        loss, batch_len = loss_func(model, d, device != "cpu")
        seen += batch_len
        nb += 1
        grads = torch.autograd.grad(loss, params)
        assert len(grads) == len(names) == len(params)
        for g_acc, g in zip(G, grads):
            ops.grad_accum(g_acc, g.detach(), square=square)  # G += |g| (g^2): ecf_grad_accum
    return G, nb


def device_get_mask_(params, grads, n_batches, mode, p, max_sparsity_per_layer, segmented=False):
    """get_mask (global) / get_layerwise_mask (``segmented``) + ``v.data *= mask`` in one go, in place on ``params``.
    Returns the per-tensor count of elements at or below the threshold."""
    datas = [q.data if q.data.is_contiguous() else q.data.contiguous() for q in params]
    table = ops.GlobalTable(datas, grads, n_batches, mode)
    protect = None
    if not segmented:
        num_to_set = [int(n * (1 - max_sparsity_per_layer)) for n in table.numels]
        if any(k > 0 for k in num_to_set):
            # threshold_t = topk(score_t, num_to_set, largest=True)[-1]; scores >= threshold_t become finfo.max
            ranks = [n - k if k > 0 else 0 for n, k in zip(table.numels, num_to_set)]
            protect = table.select(ranks, segmented=True)
            none = torch.tensor([k == 0 for k in num_to_set], device=protect.device)
            protect = torch.where(none, torch.full_like(protect, -1), protect)  # -1 = 0xffffffff: nothing protected
        num_to_zero = int(p * sum(table.numels))
        if num_to_zero == 0:
            raise IndexError("index -1 is out of bounds for dimension 0 with size 0")  # threshold[-1] of an empty top-k
        tkeys = table.select([num_to_zero - 1], segmented=False, protect=protect)
    else:
        ranks = []
        for n in table.numels:
            k = int(p * n)
            if k == 0:
                raise IndexError("index -1 is out of bounds for dimension 0 with size 0")
            ranks.append(k - 1)
        tkeys = table.select(ranks, segmented=True)
    pruned = table.apply(tkeys, segmented=segmented, protect=protect)
    for q, d in zip(params, datas):
        if d is not q.data:
            q.data.copy_(d)
    return pruned


class BLIPT5GlobalPruner(LayerWiseBasePruner):
    score_mode = None  # "mag" | "grad_mag_abs" | None (host scores)

    def __init__(self, model, data_loader, t5_prune_spec=None, vit_prune_spec=None, t5_pruning_method=None,
                 vit_pruning_method=None, t5_importance_scores_cache=None, t5_keep_indices_or_masks_cache=None,
                 vit_importance_scores_cache=None, vit_keep_indices_or_masks_cache=None, importance_scores_cache=None,
                 keep_indices_or_masks_cache=None, is_strct_pruning=False, num_samples=64, is_global=False,
                 t5_model_prefix="t5_model", vit_model_prefix="visual_encoder", sparsity_ratio_granularity=None,
                 max_sparsity_per_layer=0.8, score_method="GradMagSquare_avg", num_data_first_stage=128, num_noise=1,
                 sparsity_dict=None, prune_per_model=False, iteration=1, **kwargs):
        super().__init__(model=model, data_loader=data_loader, prune_spec=None, is_strct_pruning=is_strct_pruning,
                         importance_scores_cache=importance_scores_cache,
                         keep_indices_or_masks_cache=keep_indices_or_masks_cache, is_global=is_global,
                         num_samples=num_samples, model_prefix="tmp", sparsity_ratio_granularity=sparsity_ratio_granularity,
                         max_sparsity_per_layer=max_sparsity_per_layer, score_method=score_method,
                         num_data_first_stage=num_data_first_stage, num_noise=num_noise, sparsity_dict=sparsity_dict)
        self.t5_prune_spec = t5_prune_spec
        self.vit_prune_spec = vit_prune_spec
        self.t5_model_prefix = t5_model_prefix
        self.vit_model_prefix = vit_model_prefix
        self.prune_per_model = prune_per_model
        self.iteration = iteration

    def forward_to_cache(self, model, batch, device):
        return model(batch)

    # -- per-element scores on the device: (grads, n_batches) for the table; subclasses with host scores override _round
    def _device_scores(self, names, params):
        return None, 1

    def _round(self, names, params, p_i, max_sparsity_per_layer, masks_prev):
        grads, nb = self._device_scores(names, params)
        if self.is_global and not self.prune_per_model:
            print("global")
            device_get_mask_(params, grads, nb, self.score_mode, p_i, max_sparsity_per_layer)
        elif self.is_global and self.prune_per_model:
            print("model-level global")
            for prefix in (self.vit_model_prefix, self.t5_model_prefix):
                idx = [i for i, k in enumerate(names) if k.startswith(prefix)]
                device_get_mask_([params[i] for i in idx], None if grads is None else [grads[i] for i in idx], nb,
                                 self.score_mode, p_i, max_sparsity_per_layer)
        else:
            print("layer-wise")
            device_get_mask_(params, grads, nb, self.score_mode, p_i, max_sparsity_per_layer, segmented=True)
        return None

    def global_iterative_pruning(self, target_sparsity, dict_layers_to_prune, iteratation=1, max_sparsity_per_layer=1.0):
        names = [k for k, _ in self.model.named_parameters() if k in dict_layers_to_prune]
        params = [v for k, v in self.model.named_parameters() if k in dict_layers_to_prune]
        masks = None
        for i in range(1, iteratation + 1):
            p_i = target_sparsity ** (iteratation / i)  # modified sparsity of the i-th iteration (:166)
            masks = self._round(names, params, p_i, max_sparsity_per_layer, masks)
            print(f"Step {i}, target sparsity: {p_i:.4f}")
        for k, v in self.model.named_parameters():
            print(k, " sparsity: ", float(ops.count_zero(v.data).item()) / v.numel())
        return self.model

    @print_time
    def prune(self, importance_scores=None, keep_indices_or_masks=None):
        print("In: ", self.pruner_name)
        dtype_record, requires_grad_record, device = self.model_setup_and_record_attributes(self.model)
        if self.t5_prune_spec is None or self.vit_prune_spec is None:
            return self.model, None
        _, vit_keep, _, _ = self.convert_spec_to_list(self.vit_prune_spec)
        _, t5_keep, _, _ = self.convert_spec_to_list(self.t5_prune_spec)
        assert vit_keep == t5_keep

        def check(name, v):
            return (len(v.shape) == 2 and ".block" in name and "relative_attention_bias.weight" not in name
                    and (name.startswith(self.t5_model_prefix) or name.startswith(self.vit_model_prefix)))

        to_prune = {k: v for k, v in self.model.named_parameters() if check(k, v)}
        self.model = self.global_iterative_pruning(1 - vit_keep, to_prune, iteratation=self.iteration,
                                                   max_sparsity_per_layer=1.0)
        self.model_reset(self.model, dtype_record, requires_grad_record, device)
        return self.model, None


@registry.register_pruner("blipt5_global_mag_pruner")
class BLIPT5GlobalMagPruner(BLIPT5GlobalPruner):
    """score = float(w), signed (global_pruner.py:246-251)."""

    pruner_name = "blipt5_global_mag_pruner"
    score_mode = "mag"


@registry.register_pruner("blipt5_global_gradmagabs_pruner")
class BLIPT5GlobalGradMagAbsPruner(BLIPT5GlobalPruner):
    """score = |w| * | mean over batches of |dL/dW| |  (global_pruner.py:254-300); the gradient sums stay on the device."""

    pruner_name = "blipt5_global_gradmagabs_pruner"
    score_mode = "grad_mag_abs"

    @print_time
    def _device_scores(self, names, params):
        return accumulate_abs_grads(self.model, self.data_loader, loss_vision_language, names, params, self.num_samples)


@registry.register_pruner("blipt5_global_mezo_pruner")
class BLIPT5GlobalMeZoPruner(BLIPT5GlobalPruner):
    """One zeroth-order score per tensor (global_pruner.py:303-388): sum over the batches of |l+ - l-| / (2 eps) with
    eps = 1e-3 and ``num_noise`` draws per batch.  The mask therefore has one entry per tensor."""

    pruner_name = "blipt5_global_mezo_pruner"

    def _host_scores(self, names, params):
        from ..layer_sparsity import LayerSparsity

        model = self.model
        model.eval()
        device = next(iter(model.parameters())).device
        helper = LayerSparsity.__new__(LayerSparsity)  # only for zo_perturb_parameters (ecf_zo_perturb)
        eps = 1e-3
        scores = {}
        for i, (name, param) in enumerate(zip(names, params)):
            print(i, name)
            seen, total = 0, torch.zeros(1)
            for d in self.data_loader:
                if seen >= self.num_samples:
                    break
                per = 0.0
                for _ in range(self.num_noise):
                    if seen >= self.num_samples:
                        break
                    seed = np.random.randint(1000000000)
                    helper.zo_perturb_parameters([param], random_seed=seed, scaling_factor=1, zo_eps=eps)
                    with torch.no_grad():
                        l1, batch_len = loss_vision_language(model, d, device != "cpu")
                    helper.zo_perturb_parameters([param], random_seed=seed, scaling_factor=-2, zo_eps=eps)
                    with torch.no_grad():
                        l2, batch_len = loss_vision_language(model, d, device != "cpu")
                    helper.zo_perturb_parameters([param], random_seed=seed, scaling_factor=1, zo_eps=eps)
                    seen += batch_len
                    per += abs(((l1 - l2) / (2 * eps)).item())
                    torch.manual_seed(seed)
                total = total + torch.FloatTensor([per]).abs()
            scores[name] = total.abs()
        return scores

    def _round(self, names, params, p_i, max_sparsity_per_layer, masks_prev):
        scores = self._host_scores(names, params)
        if masks_prev is not None:
            for k in scores:
                scores[k] = scores[k] * masks_prev[k]
        masks = host_get_mask(scores, p_i, max_sparsity_per_layer) if self.is_global else host_get_layerwise_mask(scores, p_i)
        print("global" if self.is_global else "layer-wise")
        for k, v in zip(names, params):
            v.data.mul_(masks[k].to(dtype=v.dtype, device=v.device))
        return masks


def host_get_mask(importance_scores, p, max_sparsity_per_layer):
    """get_mask for a handful of host scalars (the zeroth-order variant: one score per tensor), the reference's torch
    calls verbatim in meaning (global_pruner.py:116-144)."""
    for k, v in importance_scores.items():
        num_to_set = int(v.numel() * (1 - max_sparsity_per_layer))
        if num_to_set > 0:
            thr = torch.topk(v.flatten(), num_to_set, largest=True)[0][-1]
            v[torch.where(v >= thr)] = torch.finfo(v.dtype).max
    flat = torch.cat([t.flatten() for t in importance_scores.values()])
    thr = torch.topk(flat, int(p * flat.numel()), largest=False)[0][-1]
    return {k: (v > thr).type(v.dtype) for k, v in importance_scores.items()}


def host_get_layerwise_mask(importance_scores, p):
    masks = {}
    for k, v in importance_scores.items():
        flat = v.flatten()
        thr = torch.topk(flat, int(p * flat.numel()), largest=False)[0][-1]
        masks[k] = (v > thr).type(v.dtype)
    return masks


__all__ = ["BLIPT5GlobalPruner", "BLIPT5GlobalMagPruner", "BLIPT5GlobalGradMagAbsPruner", "BLIPT5GlobalMeZoPruner",
           "device_get_mask_", "accumulate_abs_grads"]
