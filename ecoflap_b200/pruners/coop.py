"""CoOp (CLIP) entry points: ``TransformerLayer{Wanda,SparseGPT}Pruner`` and ``CLIPLayer{Wanda,SparseGPT}Pruner``
(CoOp/trainers/pruners/wanda_pruner.py:175-680, sparsegpt_pruner.py:314-805).

CLIP residual blocks use ``nn.MultiheadAttention`` whose in/out projections are applied functionally, so module
hooks never see their inputs.  The reference attaches a throw-away ``Attention`` module ("hacky_attn") that
shares the projection weights and is run once more per batch just to make the hooks fire; ``_AttentionProbe``
below plays the same role (same names ``hacky_attn.qkv`` / ``hacky_attn.proj``, same NLD layout, and -- like
the reference's shim -- no attention mask), so norms/Hessians and sparsity keys are identical.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..layer_sparsity import LayerSparsity
from . import sweep
from .base import LayerWiseBasePruner, print_time


class _AttentionProbe(nn.Module):
    """qkv / proj Linears aliasing an nn.MultiheadAttention's parameters (inputs: [N, L, D])."""

    def __init__(self, mha: nn.MultiheadAttention):
        super().__init__()
        d = mha.embed_dim
        self.num_heads = mha.num_heads
        self.qkv = nn.Linear(d, 3 * d, bias=True, device="meta")
        self.proj = nn.Linear(d, d, bias=True, device="meta")
        self.qkv.weight = nn.Parameter(mha.in_proj_weight.data, requires_grad=False)
        self.qkv.bias = nn.Parameter(mha.in_proj_bias.data, requires_grad=False)
        self.proj.weight = nn.Parameter(mha.out_proj.weight.data, requires_grad=False)
        self.proj.bias = nn.Parameter(mha.out_proj.bias.data, requires_grad=False)

    def forward(self, x):
        B, N, _ = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, -1).permute(2, 0, 3, 1, 4)
        ctx = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])  # softmax(q k^T / sqrt(d)) v, no mask
        return self.proj(ctx.transpose(1, 2).reshape(B, N, -1))


def _install_probe(layer, device):
    layer.hacky_attn = _AttentionProbe(layer.attn)
    original = layer.attention

    def attention_with_probe(x):
        layer.hacky_attn(x.permute(1, 0, 2))  # only so that the Linear hooks fire; output discarded
        return original(x)

    layer.attention = attention_with_probe

    def restore():
        # fasterprune rebinds weight.data, so hand the (pruned) tensors back to the real attention module
        layer.attn.in_proj_weight.data = layer.hacky_attn.qkv.weight.data
        layer.attn.out_proj.weight.data = layer.hacky_attn.proj.weight.data
        del layer.hacky_attn
        del layer.attention  # drop the instance attribute -> the class method is visible again

    return restore


def _clip_key(module_to_process, i, name):
    if name == "hacky_attn.qkv":
        return f"{module_to_process}.{i}.attn.in_proj_weight"
    if name == "hacky_attn.proj":
        return f"{module_to_process}.{i}.attn.out_proj.weight"
    return f"{module_to_process}.{i}.{name}.weight"


def _clip_spec():
    return sweep.SweepSpec(
        select="row",
        batch_len=lambda batch: batch["img"].shape[0],
        sample_dim=1,
        expected_nsamples=lambda inps: len(inps) * inps[0].shape[1],  # LND: batch entries are dim 1
        block_adapter=_install_probe,
        sparsity_key=_clip_key,
    )


class _ClipTransformerBase(LayerWiseBasePruner):
    method = "wanda"

    def __init__(self, model, data_loader, prune_spec=None, importance_scores_cache=None,
                 keep_indices_or_masks_cache=None, is_strct_pruning=False, num_samples=64, is_global=False,
                 model_prefix="visual", sparsity_ratio_granularity=None, max_sparsity_per_layer=0.8,
                 score_method="GradMagSquare_avg", num_data_first_stage=128, num_noise=1, sparsity_dict=None,
                 noise_eps=1e-3, **kwargs):
        super().__init__(model=model, data_loader=data_loader, prune_spec=prune_spec, is_strct_pruning=is_strct_pruning,
                         importance_scores_cache=importance_scores_cache,
                         keep_indices_or_masks_cache=keep_indices_or_masks_cache, is_global=is_global,
                         num_samples=num_samples, model_prefix=model_prefix,
                         sparsity_ratio_granularity=sparsity_ratio_granularity,
                         max_sparsity_per_layer=max_sparsity_per_layer, score_method=score_method,
                         num_data_first_stage=num_data_first_stage, num_noise=num_noise, sparsity_dict=sparsity_dict,
                         noise_eps=noise_eps)

    def reweighting_after_pruning(self, original_weights, keep_masks):
        raise NotImplementedError

    def read_cache(self, cache_file):
        raise NotImplementedError

    def check_sparsity(self, model, module_to_process="encoder.block"):
        return sweep.check_sparsity(model, module_to_process)

    def forward_to_cache(self, model, batch, device=None):
        return model.encode_image(batch["image"])

    def _call_forward_to_cache(self, model, batch, device):
        return self.forward_to_cache(model, batch, device)

    def prepare_calibration_input_encoder(self, model, dataloader, device, model_prefix, n_samples,
                                          module_to_process="encoder.block"):
        return sweep.capture_block_inputs(self, model, dataloader, device, _clip_spec(), model_prefix, n_samples,
                                          module_to_process)

    @print_time
    def _prune(self, model, dataloader, device, model_prefix, module_to_process="encoder.block", n_samples=64,
               sparsity_ratio=0.5):
        return sweep.sweep_blocks(self, model, dataloader, device, _clip_spec(), model_prefix, module_to_process,
                                  n_samples, sparsity_ratio, method=self.method)

    def get_sparsity(self, original_sparsity, sparsity_ratio_granularity=None):
        if self.sparsity_dict is not None:
            return self._load_sparsity_yaml(self.sparsity_dict)
        if sparsity_ratio_granularity is None:
            mapping = {}
        else:
            # NB the reference tests for ".blocks", which CLIP's ".resblocks" names never contain
            def accept(name, v):
                return len(v.shape) == 2 and ".blocks" in name and name.startswith(self.model_prefix)
            if sparsity_ratio_granularity == "layer":
                group = lambda k: k  # noqa: E731
            elif sparsity_ratio_granularity == "block":
                group = lambda k: ".".join(k.split(".")[:3])  # noqa: E731
            else:
                raise NotImplementedError
            mapping = {k: group(k) for k, v in self.model.named_parameters() if accept(k, v)}
        return LayerSparsity(self.model, self.data_loader, self.forward_to_cache, self.num_data_first_stage,
                             original_sparsity, self.max_sparsity_per_layer, self.score_method, self.num_noise,
                             self.noise_eps, mapping).return_sparsity()

    @print_time
    def prune(self, importance_scores=None, keep_indices_or_masks=None):
        print("In: ", self.pruner_name)
        dtype_record, requires_grad_record, device = self.model_setup_and_record_attributes(self.model)
        if self.prune_spec is None:
            return self.model, None
        _, keep_ratio, _, _ = self.convert_spec_to_list(self.prune_spec)
        sparsity_dict = self.get_sparsity(1 - keep_ratio, sparsity_ratio_granularity=self.sparsity_ratio_granularity)
        self.model = self._prune(self.model, self.data_loader, device, model_prefix=self.model_prefix,
                                 module_to_process=f"{self.model_prefix}.blocks", n_samples=self.num_samples,
                                 sparsity_ratio=sparsity_dict)
        self.model_reset(self.model, dtype_record, requires_grad_record, device)
        return self.model, sparsity_dict


class TransformerLayerWandaPruner(_ClipTransformerBase):
    pruner_name = "transformer_wanda_pruner"
    method = "wanda"


class TransformerLayerSparseGPTPruner(_ClipTransformerBase):
    pruner_name = "transformer_sparsegpt_pruner"
    method = "sparsegpt"


class _ClipBase(_ClipTransformerBase):
    def __init__(self, model, data_loader, language_prune_spec=None, visual_prune_spec=None,
                 language_pruning_method=None, visual_pruning_method=None, importance_scores_cache=None,
                 keep_indices_or_masks_cache=None, is_strct_pruning=False, num_samples=64, is_global=False,
                 language_model_prefix="transformer", visual_model_prefix="visual.transformer",
                 sparsity_ratio_granularity=None, max_sparsity_per_layer=0.8, score_method="GradMagSquare_avg",
                 num_data_first_stage=128, num_noise=1, sparsity_dict=None, noise_eps=1e-3, **kwargs):
        super().__init__(model=model, data_loader=data_loader, prune_spec=None, is_strct_pruning=is_strct_pruning,
                         importance_scores_cache=importance_scores_cache,
                         keep_indices_or_masks_cache=keep_indices_or_masks_cache, is_global=is_global,
                         num_samples=num_samples, model_prefix="tmp",
                         sparsity_ratio_granularity=sparsity_ratio_granularity,
                         max_sparsity_per_layer=max_sparsity_per_layer, score_method=score_method,
                         num_data_first_stage=num_data_first_stage, num_noise=num_noise, sparsity_dict=sparsity_dict,
                         noise_eps=noise_eps)
        self.language_prune_spec = language_prune_spec
        self.visual_prune_spec = visual_prune_spec
        self.language_model_prefix = language_model_prefix
        self.visual_model_prefix = visual_model_prefix

    def forward_to_cache(self, model, batch, device):
        """Assigned by the caller after construction (CoOp/trainers/zsclip.py:73-93): returns (loss, batch_len)."""
        pass

    def get_sparsity(self, original_sparsity, sparsity_ratio_granularity=None):
        if self.sparsity_dict is not None:
            return self._load_sparsity_yaml(self.sparsity_dict)
        lp, vp = self.language_model_prefix, self.visual_model_prefix
        if sparsity_ratio_granularity is None:
            mapping = {}
        else:
            def accept(name, v):
                return (len(v.shape) == 2 and ".resblocks" in name and "relative_attention_bias.weight" not in name
                        and (name.startswith(lp) or name.startswith(vp)))

            def tower(name, lang_value, vis_value):
                if name.startswith(lp):
                    return lang_value
                if name.startswith(vp):
                    return vis_value
                return "other"

            if sparsity_ratio_granularity == "model":
                group = lambda k: tower(k, lp, vp)  # noqa: E731
            elif sparsity_ratio_granularity == "layer":
                group = lambda k: k  # noqa: E731
            elif sparsity_ratio_granularity == "block":
                group = lambda k: tower(k, ".".join(k.split(".")[:3]), ".".join(k.split(".")[:4]))  # noqa: E731
            else:
                raise NotImplementedError
            mapping = {k: group(k) for k, v in self.model.named_parameters() if accept(k, v)}
        return LayerSparsity(self.model, self.data_loader, self.forward_to_cache, self.num_data_first_stage,
                             original_sparsity, self.max_sparsity_per_layer, self.score_method, self.num_noise,
                             self.noise_eps, mapping).return_sparsity()

    @print_time
    def prune(self, importance_scores=None, keep_indices_or_masks=None):
        print("In: ", self.pruner_name)
        dtype_record, requires_grad_record, device = self.model_setup_and_record_attributes(self.model)
        global_sparsity_dict = None
        if self.sparsity_ratio_granularity is not None:
            _, vis_keep, _, _ = self.convert_spec_to_list(self.visual_prune_spec)
            _, lang_keep, _, _ = self.convert_spec_to_list(self.language_prune_spec)
            assert vis_keep == lang_keep
            global_sparsity_dict = self.get_sparsity(1 - vis_keep, sparsity_ratio_granularity=self.sparsity_ratio_granularity)

        def ratios(spec):
            _, keep_ratio, _, _ = self.convert_spec_to_list(spec)
            if global_sparsity_dict is not None:
                return global_sparsity_dict
            return self.get_sparsity(1 - keep_ratio, sparsity_ratio_granularity=None)

        for spec, prefix in ((self.visual_prune_spec, self.visual_model_prefix),
                             (self.language_prune_spec, self.language_model_prefix)):
            if spec is not None:
                self.model = self._prune(self.model, self.data_loader, device, model_prefix=prefix,
                                         module_to_process=f"{prefix}.resblocks", n_samples=self.num_samples,
                                         sparsity_ratio=ratios(spec))
        self.model_reset(self.model, dtype_record, requires_grad_record, device)
        return self.model, global_sparsity_dict


class CLIPLayerWandaPruner(_ClipBase):
    pruner_name = "clip_wanda_pruner"
    method = "wanda"


class CLIPLayerSparseGPTPruner(_ClipBase):
    pruner_name = "clip_sparsegpt_pruner"
    method = "sparsegpt"
