"""LAVIS entry points (BLIP-2 / FlanT5 / EVA-ViT): same registry names, classes and constructor kwargs as
LAVIS/lavis/compression/pruners/{wanda,sparsegpt}_pruner.py, on top of the shared CUDA sweep engine.

  t5_wanda_pruner      T5LayerWandaPruner        wanda_pruner.py:87-375      per-ROW select, bf16 autocast
  vit_wanda_pruner     VITLayerWandaPruner       wanda_pruner.py:378-657     per-LAYER threshold select
  blipt5_wanda_pruner  BLIPT5LayerWandaPruner    wanda_pruner.py:660-876     ViT then T5 encoder/decoder
  t5_sparsegpt_pruner / vit_sparsegpt_pruner / blipt5_sparsegpt_pruner      sparsegpt_pruner.py:225-963
"""
from __future__ import annotations

from functools import partial

import torch

from ..layer_sparsity import LayerSparsity
from ..registry import registry
from . import sweep
from .base import LayerWiseBasePruner, print_time
from .losses import loss_language, loss_vision, loss_vision_language

T5_CACHE_KEYS = ("attention_mask", "position_bias", "encoder_attention_mask", "encoder_decoder_position_bias",
                 "layer_head_mask", "cross_attn_layer_head_mask", "encoder_hidden_states")


def _image_batch_len(batch):
    # wanda_pruner.py:204,488 count samples with batch["image"]; text-only batches (C4) have no "image" and
    # raise KeyError in the reference -- fall back to the text field instead of failing (documented deviation).
    if "image" in batch:
        return batch["image"].shape[0]
    return len(batch["text_input"])


def _default_key(module_to_process, i, name):
    return f"{module_to_process}.{i}.{name}.weight"


def _t5_spec(model, method):
    return sweep.SweepSpec(
        select="row",
        cache_keys=T5_CACHE_KEYS,
        block_output_index=0,
        autocast=lambda: model.maybe_autocast(dtype=torch.bfloat16),
        batch_len=_image_batch_len,
        count_by_batches=(method == "sparsegpt"),
        toggles_use_cache=True,
        # Wanda asserts len(inps) * batch size (:258); SparseGPT asserts len(inps), i.e. batch size 1 (:390)
        expected_nsamples=(lambda inps: len(inps) * inps[0].shape[0]) if method == "wanda" else (lambda inps: len(inps)),
        sparsity_key=_default_key,
    )


def _vit_spec(model, method):
    return sweep.SweepSpec(
        select="layer",
        positional_cache=("rel_pos_bias",),
        autocast=lambda: model.maybe_autocast(),
        batch_len=_image_batch_len,
        count_by_batches=(method == "sparsegpt"),
        expected_nsamples=(lambda inps: len(inps) * inps[0].shape[0]) if method == "wanda" else (lambda inps: len(inps)),
        sparsity_key=_default_key,
    )


def _group_mapping(model, accept, group_of):
    return {k: group_of(k) for k, v in model.named_parameters() if accept(k, v)}


class _LavisLayerPruner(LayerWiseBasePruner):
    """Shared plumbing of the six LAVIS classes."""

    method = "wanda"
    family = "t5"
    loss_func = None

    def reweighting_after_pruning(self, original_weights, keep_masks):
        raise NotImplementedError

    def read_cache(self, cache_file):
        raise NotImplementedError

    def _call_forward_to_cache(self, model, batch, device):
        return self.forward_to_cache(model, batch)

    def check_sparsity(self, model, module_to_process="encoder.block"):
        return sweep.check_sparsity(model, module_to_process)

    # --- family specific pieces, selected by the concrete class or (BLIP-2) per tower --------------
    def _spec(self, family=None):
        family = family or self.family
        return _t5_spec(self.model, self.method) if family == "t5" else _vit_spec(self.model, self.method)

    def _tower_family(self, module_to_process):
        return "vit" if module_to_process.endswith(".blocks") else "t5"

    def prepare_calibration_input_encoder(self, model, dataloader, device, model_prefix, n_samples,
                                          module_to_process="encoder.block"):
        spec = self._spec(self._tower_family(module_to_process))
        return sweep.capture_block_inputs(self, model, dataloader, device, spec, model_prefix, n_samples, module_to_process)

    @print_time
    def _prune(self, model, dataloader, device, model_prefix, module_to_process="encoder.block", n_samples=64,
               sparsity_ratio=0.5):
        spec = self._spec(self._tower_family(module_to_process))
        return sweep.sweep_blocks(self, model, dataloader, device, spec, model_prefix, module_to_process, n_samples,
                                  sparsity_ratio, method=self.method)

    def _layer_sparsity(self, loss, mapping, **extra):
        return LayerSparsity(self.model, self.data_loader, loss, self.num_data_first_stage, self._original_sparsity,
                             self.max_sparsity_per_layer, self.score_method, self.num_noise, self.noise_eps, mapping,
                             **extra)


# ----------------------------------------------------------------------------------------------- T5
class _T5Mixin:
    family = "t5"

    def forward_to_cache(self, model, batch):
        return model(batch)

    def get_sparsity(self, original_sparsity, sparsity_ratio_granularity=None):
        if self.sparsity_dict is not None:
            return self._load_sparsity_yaml(self.sparsity_dict)
        if sparsity_ratio_granularity is None:
            mapping = {}
        else:
            def accept(name, v):
                return (len(v.shape) == 2 and ".block" in name and "relative_attention_bias.weight" not in name
                        and name.startswith(self.model_prefix))
            if sparsity_ratio_granularity == "layer":
                mapping = _group_mapping(self.model, accept, lambda k: k)
            elif sparsity_ratio_granularity == "block":
                mapping = _group_mapping(self.model, accept, lambda k: ".".join(k.split(".")[:4]))
            else:
                raise NotImplementedError
        self._original_sparsity = original_sparsity
        extra = {"prune_per_model": self.prune_per_model} if self.method == "wanda" else {}
        return self._layer_sparsity(loss_language, mapping, **extra).return_sparsity()

    @print_time
    def prune(self, importance_scores=None, keep_indices_or_masks=None):
        print("In: ", self.pruner_name)
        dtype_record, requires_grad_record, device = self.model_setup_and_record_attributes(self.model)
        if self.prune_spec is None:
            return self.model, None
        _, keep_ratio, _, _ = self.convert_spec_to_list(self.prune_spec)
        sparsity_dict = self.get_sparsity(1 - keep_ratio, sparsity_ratio_granularity=self.sparsity_ratio_granularity)
        for tower in ("encoder", "decoder"):
            self.model = self._prune(self.model, self.data_loader, device, model_prefix=self.model_prefix,
                                     module_to_process=f"{self.model_prefix}.{tower}.block",
                                     n_samples=self.num_samples, sparsity_ratio=sparsity_dict)
        self.model_reset(self.model, dtype_record, requires_grad_record, device)
        return self.model, sparsity_dict


# ----------------------------------------------------------------------------------------------- ViT
class _VitMixin:
    family = "vit"

    def forward_to_cache(self, model, batch):
        return model.encode_image(batch["image"])

    def get_sparsity(self, original_sparsity, sparsity_ratio_granularity=None):
        if self.sparsity_dict is not None:
            sd = self._load_sparsity_yaml(self.sparsity_dict)
            # ratios produced by a BLIP-2 run use the "visual_encoder." prefix and stop at block 38 (:576-585)
            sd = {k.replace("visual_encoder.", "visual."): v for k, v in sd.items()}
            if "visual.blocks.39.attn.qkv.weight" not in sd:
                for leaf in ("attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2"):
                    sd[f"visual.blocks.39.{leaf}.weight"] = 0
            return sd
        if sparsity_ratio_granularity is None:
            mapping = {}
        else:
            def accept(name, v):
                return len(v.shape) == 2 and ".blocks" in name and name.startswith(self.model_prefix)
            if sparsity_ratio_granularity == "layer":
                mapping = _group_mapping(self.model, accept, lambda k: k)
            elif sparsity_ratio_granularity == "block":
                mapping = _group_mapping(self.model, accept, lambda k: ".".join(k.split(".")[:3]))
            else:
                raise NotImplementedError
        self._original_sparsity = original_sparsity
        extra = {"prune_per_model": self.prune_per_model} if self.method == "wanda" else {}
        return self._layer_sparsity(loss_vision, mapping, **extra).return_sparsity()

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


# ----------------------------------------------------------------------------------------------- BLIP-2
class _BlipT5Mixin:
    family = "blip"

    def _init_towers(self, t5_prune_spec, vit_prune_spec, t5_pruning_method, vit_pruning_method, t5_model_prefix,
                     vit_model_prefix):
        self.t5_prune_spec = t5_prune_spec
        self.vit_prune_spec = vit_prune_spec
        assert t5_pruning_method is not None
        assert vit_pruning_method is not None
        self.t5_model_prefix = t5_model_prefix
        self.vit_model_prefix = vit_model_prefix

    def forward_to_cache(self, model, batch):
        return model(batch)

    def get_sparsity(self, original_sparsity, sparsity_ratio_granularity=None):
        if self.sparsity_dict is not None:
            return self._load_sparsity_yaml(self.sparsity_dict)
        t5p, vitp = self.t5_model_prefix, self.vit_model_prefix
        if sparsity_ratio_granularity is None:
            mapping = {}
        else:
            def accept(name, v):
                return (len(v.shape) == 2 and ".block" in name and "relative_attention_bias.weight" not in name
                        and (name.startswith(t5p) or name.startswith(vitp)))

            def tower(name, t5_value, vit_value):
                if name.startswith(t5p):
                    return t5_value
                if name.startswith(vitp):
                    return vit_value
                return "other"

            if sparsity_ratio_granularity == "model":
                mapping = _group_mapping(self.model, accept, lambda k: tower(k, t5p, vitp))
            elif sparsity_ratio_granularity == "layer":
                mapping = _group_mapping(self.model, accept, lambda k: k)
            elif sparsity_ratio_granularity == "block":
                mapping = _group_mapping(self.model, accept, lambda k: tower(
                    k, ".".join(k.split(".")[:4]), ".".join(k.split(".")[:3])))
            else:
                raise NotImplementedError
        self._original_sparsity = original_sparsity
        extra = {"prune_per_model": self.prune_per_model, "per_model_group": [t5p, vitp]} if self.method == "wanda" else {}
        return self._layer_sparsity(loss_vision_language, mapping, **extra).return_sparsity()

    @print_time
    def prune(self, importance_scores=None, keep_indices_or_masks=None):
        print("In: ", self.pruner_name)
        dtype_record, requires_grad_record, device = self.model_setup_and_record_attributes(self.model)
        global_sparsity_dict = None
        if self.sparsity_ratio_granularity is not None:
            _, vit_keep, _, _ = self.convert_spec_to_list(self.vit_prune_spec)
            _, t5_keep, _, _ = self.convert_spec_to_list(self.t5_prune_spec)
            assert vit_keep == t5_keep
            global_sparsity_dict = self.get_sparsity(1 - vit_keep, sparsity_ratio_granularity=self.sparsity_ratio_granularity)

        def ratios(spec):
            _, keep_ratio, _, _ = self.convert_spec_to_list(spec)
            if global_sparsity_dict is not None:
                return global_sparsity_dict
            return self.get_sparsity(1 - keep_ratio, sparsity_ratio_granularity=None)

        self._frozen_towers = []  # sweep.py: swept towers answer the later capture passes from their final outputs
        if self.vit_prune_spec is not None:
            self.model = self._prune(self.model, self.data_loader, device, model_prefix=self.vit_model_prefix,
                                     module_to_process=f"{self.vit_model_prefix}.blocks", n_samples=self.num_samples,
                                     sparsity_ratio=ratios(self.vit_prune_spec))
        if self.t5_prune_spec is not None:
            sd = ratios(self.t5_prune_spec)
            for tower in ("encoder", "decoder"):
                self.model = self._prune(self.model, self.data_loader, device, model_prefix=self.t5_model_prefix,
                                         module_to_process=f"{self.t5_model_prefix}.{tower}.block",
                                         n_samples=self.num_samples, sparsity_ratio=sd)
        self._frozen_towers = None
        self.model_reset(self.model, dtype_record, requires_grad_record, device)
        return self.model, global_sparsity_dict


def _single_tower_init(self, model, data_loader, default_prefix, kw):
    kw.setdefault("model_prefix", default_prefix)
    LayerWiseBasePruner.__init__(self, model=model, data_loader=data_loader, **kw)


def _make_single(name, cls_name, mixin, method, default_prefix, loss):
    def __init__(self, model, data_loader, prune_spec=None, importance_scores_cache=None,
                 keep_indices_or_masks_cache=None, is_strct_pruning=False, num_samples=64, is_global=False,
                 model_prefix=default_prefix, sparsity_ratio_granularity=None, max_sparsity_per_layer=0.8,
                 score_method="GradMagSquare_avg", num_data_first_stage=128, num_noise=1, sparsity_dict=None,
                 noise_eps=1e-3, prune_per_model=False, **kwargs):
        LayerWiseBasePruner.__init__(
            self, model=model, data_loader=data_loader, prune_spec=prune_spec, is_strct_pruning=is_strct_pruning,
            importance_scores_cache=importance_scores_cache, keep_indices_or_masks_cache=keep_indices_or_masks_cache,
            is_global=is_global, num_samples=num_samples, model_prefix=model_prefix,
            sparsity_ratio_granularity=sparsity_ratio_granularity, max_sparsity_per_layer=max_sparsity_per_layer,
            score_method=score_method, num_data_first_stage=num_data_first_stage, num_noise=num_noise,
            sparsity_dict=sparsity_dict, noise_eps=noise_eps, prune_per_model=prune_per_model)
        self.loss_func = loss

    cls = type(cls_name, (mixin, _LavisLayerPruner), {"__init__": __init__, "pruner_name": name, "method": method})
    return registry.register_pruner(name)(cls)


def _make_blip(name, cls_name, method):
    def __init__(self, model, data_loader, t5_prune_spec=None, vit_prune_spec=None, t5_pruning_method=None,
                 vit_pruning_method=None, t5_importance_scores_cache=None, t5_keep_indices_or_masks_cache=None,
                 vit_importance_scores_cache=None, vit_keep_indices_or_masks_cache=None, importance_scores_cache=None,
                 keep_indices_or_masks_cache=None, is_strct_pruning=False, num_samples=64, is_global=False,
                 t5_model_prefix="t5_model", vit_model_prefix="visual_encoder", sparsity_ratio_granularity=None,
                 max_sparsity_per_layer=0.8, score_method="GradMagSquare_avg", num_data_first_stage=128, num_noise=1,
                 sparsity_dict=None, noise_eps=1e-3, prune_per_model=False, **kwargs):
        LayerWiseBasePruner.__init__(
            self, model=model, data_loader=data_loader, prune_spec=None, is_strct_pruning=is_strct_pruning,
            importance_scores_cache=importance_scores_cache, keep_indices_or_masks_cache=keep_indices_or_masks_cache,
            is_global=is_global, num_samples=num_samples, model_prefix="tmp",
            sparsity_ratio_granularity=sparsity_ratio_granularity, max_sparsity_per_layer=max_sparsity_per_layer,
            score_method=score_method, num_data_first_stage=num_data_first_stage, num_noise=num_noise,
            sparsity_dict=sparsity_dict, noise_eps=noise_eps, prune_per_model=prune_per_model)
        self._init_towers(t5_prune_spec, vit_prune_spec, t5_pruning_method, vit_pruning_method, t5_model_prefix,
                          vit_model_prefix)

    cls = type(cls_name, (_BlipT5Mixin, _LavisLayerPruner), {"__init__": __init__, "pruner_name": name, "method": method})
    return registry.register_pruner(name)(cls)


T5LayerWandaPruner = _make_single("t5_wanda_pruner", "T5LayerWandaPruner", _T5Mixin, "wanda", "t5_model", loss_language)
VITLayerWandaPruner = _make_single("vit_wanda_pruner", "VITLayerWandaPruner", _VitMixin, "wanda", "visual", loss_vision)
BLIPT5LayerWandaPruner = _make_blip("blipt5_wanda_pruner", "BLIPT5LayerWandaPruner", "wanda")
T5LayerSparseGPTPruner = _make_single("t5_sparsegpt_pruner", "T5LayerSparseGPTPruner", _T5Mixin, "sparsegpt", "t5_model", loss_language)
VITLayerSparseGPTPruner = _make_single("vit_sparsegpt_pruner", "VITLayerSparseGPTPruner", _VitMixin, "sparsegpt", "visual", loss_vision)
BLIPT5LayerSparseGPTPruner = _make_blip("blipt5_sparsegpt_pruner", "BLIPT5LayerSparseGPTPruner", "sparsegpt")

__all__ = ["T5LayerWandaPruner", "VITLayerWandaPruner", "BLIPT5LayerWandaPruner", "T5LayerSparseGPTPruner",
           "VITLayerSparseGPTPruner", "BLIPT5LayerSparseGPTPruner", "partial"]
