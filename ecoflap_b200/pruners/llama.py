"""LLaMA entry points.  The reference imports ``prune_wanda / prune_sparsegpt / prune_magnitude /
check_sparsity`` from ``LLaMA/lib`` (LLaMA/main.py:8-9), but that directory was never committed
(``.gitignore:17`` ignores ``lib/``), so PARITY IS UNPINNED for this family: the functions below follow the
contract visible in ``LLaMA/main.py:31-80`` and ``LLaMA/scripts/ecoflap_{zero,first}.sh`` (flags
``--approach_for_sparsity --aggregate_method --score_method --use_mezo --num_samples_for_first_stage
--max_sparsity_per_layer``), the upstream-Wanda per-row selection (``LLaMA/image_classifiers/prune_utils.py:35-38``)
and the LAVIS row-variant sweep (``wanda_pruner.py:217-290``), whose kernels they share.

Duck type: ``model.model.layers`` is the decoder stack; ``model(input_ids, labels=input_ids).loss`` is the LM loss;
calibration batches are ``(input_ids, targets)`` tuples of shape [1, seqlen] as produced by upstream ``get_loaders``.
"""
from __future__ import annotations

import torch

from ..layer_sparsity import LayerSparsity
from . import sweep
from .base import LayerWiseBasePruner

LLAMA_CACHE_KEYS = ("attention_mask", "position_ids", "position_embeddings")


class _LlamaSweepPruner(LayerWiseBasePruner):
    """Adapter that lets the shared sweep engine drive a HF-style decoder stack."""

    pruner_name = "llama_wanda_pruner"

    def __init__(self, model, data_loader, method="wanda", **kw):
        super().__init__(model=model, data_loader=data_loader, model_prefix="model", **kw)
        self.method = method

    def _call_forward_to_cache(self, model, batch, device):
        return model(batch[0].to(device))

    def _spec(self):
        return sweep.SweepSpec(
            select="row", cache_keys=LLAMA_CACHE_KEYS, optional_keys=True, block_output_index=0,
            batch_len=lambda batch: batch[0].shape[0],
            expected_nsamples=lambda inps: len(inps) * inps[0].shape[0],
            sparsity_key=lambda stack, i, name: f"{stack}.{i}.{name}.weight")

    def prepare_calibration_input_encoder(self, model, dataloader, device, model_prefix, n_samples, module_to_process):
        return sweep.capture_block_inputs(self, model, dataloader, device, self._spec(), model_prefix, n_samples,
                                          module_to_process)

    def _prune(self, model, dataloader, device, module_to_process, n_samples, sparsity_ratio):
        return sweep.sweep_blocks(self, model, dataloader, device, self._spec(), "model", module_to_process, n_samples,
                                  sparsity_ratio, method=self.method)


def _lm_loss(model, batch, cuda_enabled):
    dev = next(iter(model.parameters())).device
    ids = batch[0].to(dev)
    return model(ids, labels=ids).loss, ids.shape[0]


def _ratios(args, model, dataloader):
    """Uniform ratio, or the ECoFLaP allocation when --approach_for_sparsity is set."""
    approach = getattr(args, "approach_for_sparsity", None)
    if approach is None:
        return _Uniform(args.sparsity_ratio)
    compute = getattr(args, "score_method", "GradOnly")
    if getattr(args, "use_mezo", False):
        compute = "MEZO-" + compute
    method = f"{compute}_{getattr(args, 'aggregate_method', 'sum')}"

    def accept(name, v):
        return len(v.shape) == 2 and ".layers." in name

    if approach == "layer":
        group = lambda k: k  # noqa: E731
    elif approach == "block":
        group = lambda k: ".".join(k.split(".")[:3])  # model.layers.<i>  # noqa: E731
    else:
        raise NotImplementedError(approach)
    mapping = {k: group(k) for k, v in model.named_parameters() if accept(k, v)}
    ls = LayerSparsity(model, dataloader, _lm_loss, getattr(args, "num_samples_for_first_stage", 32),
                       args.sparsity_ratio, getattr(args, "max_sparsity_per_layer", 0.7), method,
                       layer_to_group_mapping=mapping)
    return ls.return_sparsity()


class _Uniform:
    def __init__(self, r):
        self.r = r

    def __getitem__(self, key):
        return self.r


def _run(args, model, dataloader, device, prune_n, prune_m, method):
    use_cache = getattr(model.config, "use_cache", None)
    if use_cache is not None:
        model.config.use_cache = False
    try:
        ratios = _ratios(args, model, dataloader)
        pruner = _LlamaSweepPruner(model, dataloader, method=method, prune_spec=None)
        pruner.prune_n, pruner.prune_m = int(prune_n), int(prune_m)  # --sparsity_type 2:4 / 4:8 (LLaMA/main.py:55-58)
        with torch.no_grad():
            pruner._prune(model, dataloader, device, "model.layers", getattr(args, "nsamples", 128), ratios)
    finally:
        if use_cache is not None:
            model.config.use_cache = use_cache
    return model


def prune_wanda(args, model, tokenizer=None, device=torch.device("cuda:0"), prune_n=0, prune_m=0, dataloader=None):
    """Wanda (optionally with ECoFLaP ratios) on a LLaMA-style decoder.  ``dataloader`` replaces upstream's
    ``get_loaders("c4", ...)`` (no dataset access here): an iterable of (input_ids, targets)."""
    return _run(args, model, dataloader, device, prune_n, prune_m, "wanda")


def prune_sparsegpt(args, model, tokenizer=None, device=torch.device("cuda:0"), prune_n=0, prune_m=0, dataloader=None):
    return _run(args, model, dataloader, device, prune_n, prune_m, "sparsegpt")


def prune_magnitude(args, model, tokenizer=None, device=torch.device("cuda:0"), prune_n=0, prune_m=0):
    """Magnitude baseline (LLaMA/main.py:8,76-77 -> upstream Wanda's lib/prune.py, absent from the reference checkout:
    PARITY UNPINNED).  Upstream: per Linear  W_metric = |W|;  thresh = sort(W_metric.flatten())[int(numel * s)];
    W[W_metric <= thresh] = 0  -- the per-layer threshold rule of A5 with a unit norm (fp32(|w|) * 1 orders and ties exactly
    like the fp16 |w|), so it is the same fused kernel with scaler_row = 1; n:m takes the n:m kernel the same way."""
    from .. import ops

    layers = model.model.layers
    for i in range(len(layers)):
        subset = sweep.find_layers(layers[i])
        items = []
        for name in subset:
            W = subset[name].weight.data
            ones = torch.ones(W.shape[1], dtype=torch.float32, device=W.device)
            if prune_n != 0:
                ops.wanda_nm_select_apply(W, ones, int(prune_n), int(prune_m))
            else:
                items.append((W, ones, int(W.numel() * args.sparsity_ratio)))
        for j in range(0, len(items), 8):  # the Linears of a block share launches (ECF_LAYER_MAX_BATCH = 8)
            ops.wanda_layer_thresh_apply_batched(items[j:j + 8])
    return model


def check_sparsity(model):
    return sweep.check_sparsity(model, "model.layers")
