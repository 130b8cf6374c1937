"""Stage 2 of every layer-wise pruner: the calibration sweep over a stack of transformer blocks.

One engine replaces the near-identical ``prepare_calibration_input_encoder`` + ``_prune`` pairs of the
reference (LAVIS wanda_pruner.py:168-290,455-568; sparsegpt_pruner.py:304-407,569-664; CoOp
wanda_pruner.py:247-401; UPop wanda_pruner.py:151-290,417-530).  The control flow is the reference's:

    capture block-0 inputs (+ kwargs) for every calibration batch by aborting the forward pass
    for each block:   hook its Linears -> forward all batches -> score/select/apply -> forward again -> swap

The per-family differences (what the Catcher records, how a block is called, which autocast is active,
which selection variant is used, how a batch is counted) are data in a ``SweepSpec``; the numerical work
is done by the CUDA kernels behind ``WrappedGPT`` / ``SparseGPT`` / ``ops``.
"""
from __future__ import annotations

import contextlib
import os
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import torch
import torch.nn as nn

from .. import dist as edist
from .. import ops
from ..accumulators import HessianBatch, NormBatch, SparseGPT, WrappedGPT


def get_module_recursive(base, module_to_process):
    """'a.b.c' -> base.a.b.c (wanda_pruner.py:19-30)."""
    for part in [p for p in module_to_process.split(".") if p != ""]:
        base = getattr(base, part)
    return base


def find_layers(module, layers=(nn.Linear,), name=""):
    """{qualified name: module} of the given exact types below ``module`` (wanda_pruner.py:33-52)."""
    if type(module) in tuple(layers):
        return {name: module}
    found = {}
    for child_name, child in module.named_children():
        found.update(find_layers(child, layers=layers, name=name + "." + child_name if name != "" else child_name))
    return found


class _StopForward(ValueError):
    """The reference aborts the capture forward by raising ValueError; same type so that user code which
    catches ValueError around forward_to_cache keeps working."""


@dataclass
class SweepSpec:
    select: str                                   # "row" | "layer"  (Wanda) -- ignored by SparseGPT
    cache_keys: Sequence[str] = ()                # kwargs recorded by the Catcher and replayed to every block
    optional_keys: bool = False                   # UPop records a key only when it is present
    positional_cache: Sequence[str] = ()          # EVA ViT passes rel_pos_bias positionally
    block_output_index: Optional[int] = None      # T5/BERT blocks return tuples -> [0]
    autocast: Optional[Callable] = None           # () -> context manager
    batch_len: Callable = None                    # batch -> number of samples (capture loop bound)
    count_by_batches: bool = False                # SparseGPT variants stop after n_samples *batches*
    sample_dim: int = 0                           # dim of inps[j] holding the batch entries (CLIP is LND -> 1)
    toggles_use_cache: bool = False               # T5/BERT: model.<prefix>.config.use_cache = False during the sweep
    expected_nsamples: Callable = None            # (inps) -> value WrappedGPT/SparseGPT.nsamples must reach
    block_adapter: Optional[Callable] = None      # CLIP: exposes nn.MultiheadAttention projections to hooks
    sparsity_key: Callable = None                 # (module_to_process, i, name) -> key into the sparsity dict


def _nullcontext():
    return contextlib.nullcontext()


class _Catcher(nn.Module):
    def __init__(self, module, spec, inps, caches):
        super().__init__()
        self.module = module
        self._spec, self._inps, self._caches = spec, inps, caches

    def forward(self, inp, *args, **kwargs):
        spec = self._spec
        self._inps.append(inp)
        self._inps[-1].requires_grad = False
        cache = {}
        for name, value in zip(spec.positional_cache, args):
            cache[name] = value
        for k in spec.positional_cache:
            if k in kwargs:
                cache[k] = kwargs[k]
        for k in spec.cache_keys:
            if spec.optional_keys:
                if k in kwargs:
                    cache[k] = kwargs[k]
            else:
                cache[k] = kwargs[k]
        self._caches.append(cache)
        raise _StopForward


def capture_block_inputs(pruner, model, dataloader, device, spec: SweepSpec, model_prefix, n_samples, module_to_process):
    """prepare_calibration_input_encoder: returns (inps, outs, caches)."""
    stem = getattr(model, model_prefix, None) if spec.toggles_use_cache else None
    if stem is not None:
        use_cache = stem.config.use_cache
        stem.config.use_cache = False
    layers = get_module_recursive(model, module_to_process)
    inps, caches = [], []
    layers[0] = _Catcher(layers[0], spec, inps, caches)
    try:
        seen = 0
        for i, batch in enumerate(dataloader):
            if spec.count_by_batches:
                if i >= n_samples:
                    break
            else:
                if seen >= n_samples:
                    break
                seen += spec.batch_len(batch)
            try:
                pruner._call_forward_to_cache(model, batch, device)
            except ValueError:
                pass
    finally:
        layers[0] = layers[0].module
        if stem is not None:
            stem.config.use_cache = use_cache
    outs = [None] * len(inps)
    return inps, outs, caches


def _run_block(layer, inp, cache, spec):
    out = layer(inp, **cache)
    # T5/BERT blocks return tuples; a block that returns a bare tensor (transformers >= 5 LlamaDecoderLayer) must not be
    # indexed -- out[0] would silently drop the batch dimension
    if spec.block_output_index is not None and isinstance(out, (tuple, list)):
        out = out[spec.block_output_index]
    return out


def wanda_prune_linear(linear, acc: WrappedGPT, sparsity, select):
    """score + select + apply for one Linear: one fused kernel call, weights zeroed in place.
    k / kth index are computed here with the reference's Python expressions (wanda_pruner.py:276, :555)."""
    W = linear.weight.data
    rows, cols = W.shape
    if select == "row":
        ops.wanda_row_select_apply(W, acc.scaler_row, int(cols * sparsity))
    elif select == "layer":
        ops.wanda_layer_thresh_apply(W, acc.scaler_row, int(W.numel() * sparsity))
    else:
        raise ValueError(select)


def sweep_blocks(pruner, model, dataloader, device, spec: SweepSpec, model_prefix, module_to_process, n_samples,
                 sparsity_ratio, method="wanda"):
    """_prune: the block-by-block calibrate / prune / re-forward loop."""
    stem = getattr(model, model_prefix, None) if spec.toggles_use_cache else None
    if stem is not None:
        use_cache = stem.config.use_cache
        stem.config.use_cache = False
    print("loading calibdation data")
    with torch.no_grad():
        inps, outs, caches = pruner.prepare_calibration_input_encoder(model, dataloader, device, model_prefix, n_samples,
                                                                     module_to_process)
    n_batches = min(n_samples, len(inps))
    # P ranks (the reference is single-GPU): the calibration batches are sharded round-robin -- rank r forwards
    # batches j = r (mod P) through every block -- the block's norms / Hessians are merged by one exchange, and every
    # rank prunes (Wanda: replicated, identical norms give identical masks; SparseGPT: Linear i on rank i mod P, then a
    # broadcast of the pruned weights).  A rank only ever reads the activations of its own batches.
    rank, world = edist.rank_world() if edist.is_dist() else (0, 1)
    my_batches = list(range(n_batches)) if world == 1 else edist.shard_indices(n_batches, rank, world)
    autocast = spec.autocast or _nullcontext
    expected_nsamples = spec.expected_nsamples(inps)  # from the catcher's complete list (later lists are per-rank sparse)
    layers = get_module_recursive(model, module_to_process)
    acc_cls = WrappedGPT if method == "wanda" else SparseGPT

    for i in range(len(layers)):
        layer = layers[i]
        restore = spec.block_adapter(layer, device) if spec.block_adapter is not None else None
        subset = find_layers(layer)
        # Wanda: the hook calls of the block's calibration sweep are deferred and become one batched norm launch
        norm_batch = NormBatch() if (method == "wanda" and os.environ.get("ECF_NORM_BATCH", "1") != "0") else None
        # SparseGPT: likewise -- one tensor-core launch per distinct hook input over the concatenated batches
        hess_batch = HessianBatch() if (method != "wanda" and os.environ.get("ECF_HESSIAN_BATCH", "1") != "0") else None
        if method == "wanda":
            wrapped = {name: WrappedGPT(subset[name], batch=norm_batch) for name in subset}
        else:
            wrapped = {name: acc_cls(subset[name], batch=hess_batch) for name in subset}

        def make_hook(name):
            permute = restore is not None and not name.startswith("hacky")

            def hook(_, inp, out):
                # detach() shares the version counter with the live activation (``.data`` has its own), so that
                # NormBatch.flush() notices an in-place update between this hook and the deferred launch
                x = inp[0].detach()
                if permute:  # CLIP blocks are LND: the MLP hooks see [L, N, D] (CoOp wanda_pruner.py:349-353)
                    x = x.permute(1, 0, 2)
                wrapped[name].add_batch(x, out.data)

            return hook

        handles = [subset[name].register_forward_hook(make_hook(name)) for name in wrapped]
        try:
            for j in my_batches:
                with torch.no_grad():
                    with autocast():
                        outs[j] = _run_block(layer, inps[j], caches[j], spec)
            if norm_batch is not None:  # one launch for the whole calibration sweep of this block
                norm_batch.flush()
            if hess_batch is not None:
                hess_batch.flush()
            if world > 1:  # global running means over the batches of all ranks (one all-reduce for the block)
                if method == "wanda":
                    edist.sync_block_norms(list(wrapped.values()))
                else:
                    edist.sync_block_hessians(list(wrapped.values()))
        finally:
            for h in handles:
                h.remove()

        layer_items, row_items, obs_owners = [], [], []
        for name in subset:
            assert wrapped[name].nsamples == expected_nsamples, (
                f"{name}: accumulated {wrapped[name].nsamples} samples, expected {expected_nsamples}")
            print(f"pruning layer {i} name {name}")
            key = spec.sparsity_key(module_to_process, i, name)
            if method == "wanda" and pruner.prune_n != 0:
                # structured n:m branch (wanda_pruner.py:265-270 / :546-551): same for the row and the layer variants
                ops.wanda_nm_select_apply(subset[name].weight.data, wrapped[name].scaler_row, pruner.prune_n, pruner.prune_m)
            elif method == "wanda" and spec.select == "layer":
                # every Linear of the block keeps its own exact threshold; they are selected in one cooperative launch
                W = subset[name].weight.data
                layer_items.append((W, wrapped[name].scaler_row, int(W.numel() * sparsity_ratio[key])))
            elif method == "wanda" and spec.select == "row":
                # k per row with the reference's Python expression (wanda_pruner.py:276); the Linears of the block that
                # share a row length are selected in one launch
                W = subset[name].weight.data
                row_items.append((W, wrapped[name].scaler_row, int(W.shape[1] * sparsity_ratio[key])))
            elif method == "wanda":
                wanda_prune_linear(subset[name], wrapped[name], sparsity_ratio[key], spec.select)
            else:
                owner = len(obs_owners) % world  # the OBS problems of a block are independent: Linear i on rank i mod P
                obs_owners.append((name, owner))
                if owner == rank:
                    wrapped[name].fasterprune(sparsity_ratio[key], prune_n=pruner.prune_n, prune_m=pruner.prune_m,
                                              percdamp=0.01, blocksize=128)
                wrapped[name].free()
        if layer_items:
            ops.wanda_layer_thresh_apply_batched(layer_items)
        if row_items:
            ops.wanda_row_select_apply_batched(row_items)
        if world > 1:
            import torch.distributed as dist

            for name, owner in obs_owners:  # every rank continues with the owner's pruned weights
                dist.broadcast(subset[name].weight.data, src=owner)
        if restore is not None:
            restore()

        for j in my_batches:
            with torch.no_grad():
                with autocast():
                    outs[j] = _run_block(layer, inps[j], caches[j], spec)
        inps, outs = outs, inps

    if stem is not None:
        stem.config.use_cache = use_cache
    if torch.cuda.is_available():
        torch.cuda.empty_cache()
    return model


def check_sparsity(model, module_to_process):
    """Fraction of zero weights in the Linears of a block stack (wanda_pruner.py:139-163,430-450), counted on
    the device with ecf_count_zero."""
    layers = get_module_recursive(model, module_to_process)
    count, total = 0, 0
    for i in range(len(layers)):
        subset = find_layers(layers[i])
        sub_count, sub_params = 0, 0
        for name in subset:
            W = subset[name].weight.data
            z = int(ops.count_zero(W).item())
            count += z
            sub_count += z
            total += W.numel()
            sub_params += W.numel()
        print(f"layer {i} sparsity {float(sub_count) / sub_params:.6f}")
    return float(count) / total
