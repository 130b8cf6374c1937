"""UPop (BLIP / BERT) entry points: ``BertLayerWandaPruner``, ``VITLayerWandaPruner`` and
``BLIPBertLayerWandaPruner(task=...)`` (UPop/pruners/wanda_pruner.py:81-834).  Batches are tuples
(``batch[0].shape[0]`` samples), blocks run without autocast, BERT layers use the per-row select and the ViT
the per-layer threshold.

Reference quirk, reproduced by default: the reference builds ``LayerSparsity(..., self.score_method, self.task,
mapping)`` positionally (wanda_pruner.py:707-717), which binds ``task`` to ``num_noise`` and the mapping to
``noise_eps`` and leaves ``layer_to_group_mapping`` empty -- ECoFLaP allocation silently degenerates to uniform
sparsity in that copy.  A drop-in must give the reference's result, so ``reference_uniform_quirk=True`` is the default;
``reference_uniform_quirk=False`` is the opt-in fix (arguments passed by keyword, the requested granularity takes
effect).
"""
from __future__ import annotations

import torch

from ..layer_sparsity import LayerSparsity
from . import sweep
from .base import LayerWiseBasePruner, print_time

BERT_CACHE_KEYS = ("attention_mask", "head_mask", "encoder_hidden_states", "encoder_attention_mask",
                   "output_attentions", "mode", "labels")


def _tuple_batch_len(batch):
    return batch[0].shape[0]


def _default_key(module_to_process, i, name):
    return f"{module_to_process}.{i}.{name}.weight"


def _bert_spec():
    return sweep.SweepSpec(select="row", cache_keys=BERT_CACHE_KEYS, optional_keys=True, block_output_index=0,
                           batch_len=_tuple_batch_len, toggles_use_cache=True,
                           expected_nsamples=lambda inps: len(inps) * inps[0].shape[0], sparsity_key=_default_key)


def _vit_spec():
    return sweep.SweepSpec(select="layer", positional_cache=("register_hook",), batch_len=_tuple_batch_len,
                           expected_nsamples=lambda inps: len(inps) * inps[0].shape[0], sparsity_key=_default_key)


class _UPopBase(LayerWiseBasePruner):
    family = "bert"

    def reweighting_after_pruning(self, original_weights, keep_masks):
        raise NotImplementedError

    def read_cache(self, cache_file):
        raise NotImplementedError

    def check_sparsity(self, model, module_to_process="encoder.block"):
        return sweep.check_sparsity(model, module_to_process)

    def _call_forward_to_cache(self, model, batch, device):
        return self.forward_to_cache(model, batch, device)

    def _spec_for(self, module_to_process):
        if not module_to_process.endswith(".blocks"):
            return _bert_spec()
        spec = _vit_spec()
        if getattr(self, "task", None) == "nlvr":
            # the reference asserts twice the count for NLVR's image pairs (UPop wanda_pruner.py:496-497); kept verbatim
            spec.expected_nsamples = lambda inps: len(inps) * inps[0].shape[0] * 2
        return spec

    def prepare_calibration_input_encoder(self, model, dataloader, device, model_prefix, n_samples,
                                          module_to_process="encoder.block"):
        return sweep.capture_block_inputs(self, model, dataloader, device, self._spec_for(module_to_process),
                                          model_prefix, n_samples, module_to_process)

    @print_time
    def _prune(self, model, dataloader, device, model_prefix, module_to_process="encoder.block", n_samples=64,
               sparsity_ratio=0.5):
        return sweep.sweep_blocks(self, model, dataloader, device, self._spec_for(module_to_process), model_prefix,
                                  module_to_process, n_samples, sparsity_ratio, method="wanda")


class BertLayerWandaPruner(_UPopBase):
    pruner_name = "bert_wanda_pruner"

    def forward_to_cache(self, model, batch, device="cuda"):
        return model(batch)


class VITLayerWandaPruner(_UPopBase):
    pruner_name = "vit_wanda_pruner"

    def forward_to_cache(self, model, batch, device="cuda"):
        return model.encode_image(batch["image"])


class BLIPBertLayerWandaPruner(_UPopBase):
    pruner_name = "blipbert_wanda_pruner"

    def __init__(self, model, data_loader, bert_prune_spec=None, vit_prune_spec=None, bert_pruning_method=None,
                 vit_pruning_method=None, bert_importance_scores_cache=None, bert_keep_indices_or_masks_cache=None,
                 vit_importance_scores_cache=None, vit_keep_indices_or_masks_cache=None, importance_scores_cache=None,
                 keep_indices_or_masks_cache=None, is_strct_pruning=False, num_samples=64, is_global=False,
                 bert_model_prefix="text_encoder", vit_model_prefix="visual_encoder", sparsity_ratio_granularity=None,
                 max_sparsity_per_layer=0.8, score_method="GradMagSquare_avg", num_data_first_stage=128, task="nlvr",
                 reference_uniform_quirk=True, **kwargs):
        super().__init__(model=model, data_loader=data_loader, prune_spec=None, is_strct_pruning=is_strct_pruning,
                         importance_scores_cache=importance_scores_cache,
                         keep_indices_or_masks_cache=keep_indices_or_masks_cache, is_global=is_global,
                         num_samples=num_samples, model_prefix="tmp",
                         sparsity_ratio_granularity=sparsity_ratio_granularity,
                         max_sparsity_per_layer=max_sparsity_per_layer, score_method=score_method,
                         num_data_first_stage=num_data_first_stage)
        self.task = task
        self.bert_prune_spec = bert_prune_spec
        self.vit_prune_spec = vit_prune_spec
        self.bert_model_prefix = bert_model_prefix
        self.vit_model_prefix = vit_model_prefix
        self.reference_uniform_quirk = reference_uniform_quirk

    def forward_to_cache(self, model, batch, device="cuda"):
        """Task-specific loss closure, returns (loss, batch_len) (wanda_pruner.py:721-748)."""
        if self.task == "nlvr":
            image0, image1, text, targets = batch
            images = torch.cat([image0, image1], dim=0).to(device)
            return model(images, text, targets=targets.to(device), train=True), image0.shape[0]
        if self.task == "coco":
            image, caption, _ = batch
            return model(image.to(device), caption), image.shape[0]
        if self.task == "retrieval":
            image, caption, idx = batch
            image, idx = image.to(device, non_blocking=True), idx.to(device, non_blocking=True)
            return model.forward_itm(image, caption, alpha=0.4, idx=idx), image.shape[0]
        if self.task == "vqa":
            image, question, answer, weights, n = batch
            image, weights = image.to(device, non_blocking=True), weights.to(device, non_blocking=True)
            return model(image, question, answer, train=True, n=n, weights=weights), image.shape[0]
        return model(batch), 1

    def get_sparsity(self, original_sparsity, sparsity_ratio_granularity=None):
        bp, vp = self.bert_model_prefix, self.vit_model_prefix
        if sparsity_ratio_granularity is None or self.reference_uniform_quirk:
            mapping = {}
        else:
            def accept(name, v):
                return (len(v.shape) == 2 and (".block" in name or ".layer" in name)
                        and "relative_attention_bias.weight" not in name
                        and (name.startswith(bp + ".") or name.startswith(vp + ".") or name.startswith("text_encoder.")))

            def block_group(name):
                if name.startswith(bp + "."):
                    return ".".join(name.split(".")[:5 if self.task in ("coco", "vqa") else 4])
                if name.startswith(vp + "."):
                    return ".".join(name.split(".")[:3])
                if name.startswith("text_encoder."):
                    return ".".join(name.split(".")[:4])
                return "other"

            if sparsity_ratio_granularity == "model":
                group = lambda k: bp if k.startswith(bp) else (vp if k.startswith(vp) else "other")  # noqa: E731
            elif sparsity_ratio_granularity == "layer":
                group = lambda k: k  # noqa: E731
            elif sparsity_ratio_granularity == "block":
                group = block_group
            else:
                raise NotImplementedError
            mapping = {k: group(k) for k, v in self.model.named_parameters() if accept(k, v)}
        return LayerSparsity(self.model, self.data_loader, self.forward_to_cache, self.num_data_first_stage,
                             original_sparsity, self.max_sparsity_per_layer, self.score_method,
                             layer_to_group_mapping=mapping).return_sparsity()

    @print_time
    def prune(self, importance_scores=None, keep_indices_or_masks=None):
        print("In: ", self.pruner_name)
        dtype_record, requires_grad_record, device = self.model_setup_and_record_attributes(self.model)
        global_sparsity_dict = None
        if self.sparsity_ratio_granularity is not None:
            _, vit_keep, _, _ = self.convert_spec_to_list(self.vit_prune_spec)
            _, bert_keep, _, _ = self.convert_spec_to_list(self.bert_prune_spec)
            assert vit_keep == bert_keep
            global_sparsity_dict = self.get_sparsity(1 - vit_keep, sparsity_ratio_granularity=self.sparsity_ratio_granularity)

        def ratios(spec):
            _, keep_ratio, _, _ = self.convert_spec_to_list(spec)
            if global_sparsity_dict is not None:
                return global_sparsity_dict
            return self.get_sparsity(1 - keep_ratio, sparsity_ratio_granularity=None)

        if self.vit_prune_spec is not None:
            self.model = self._prune(self.model, self.data_loader, device, model_prefix=self.vit_model_prefix,
                                     module_to_process=f"{self.vit_model_prefix}.blocks", n_samples=self.num_samples,
                                     sparsity_ratio=ratios(self.vit_prune_spec))
        if self.bert_prune_spec is not None and getattr(self.model, self.bert_model_prefix, None) is not None:
            sd = ratios(self.bert_prune_spec)
            if self.task == "vqa":
                self.model = self._prune(self.model, self.data_loader, device, model_prefix="text_encoder",
                                         module_to_process="text_encoder.encoder.layer", n_samples=self.num_samples,
                                         sparsity_ratio=sd)
            if self.task in ("coco", "vqa"):
                stack = f"{self.bert_model_prefix}.bert.encoder.layer"
            else:
                stack = f"{self.bert_model_prefix}.encoder.layer"
            self.model = self._prune(self.model, self.data_loader, device, model_prefix=self.bert_model_prefix,
                                     module_to_process=stack, n_samples=self.num_samples, sparsity_ratio=sd)
        self.model_reset(self.model, dtype_record, requires_grad_record, device)
        return self.model, global_sparsity_dict
