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
from .. import graphs
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
    frozen = []
    if os.environ.get("ECF_TOWER_MEMO", "1") != "0" and spec.batch_len is not None:
        for path, f_outs, f_dim in getattr(pruner, "_frozen_towers", None) or []:
            if path != module_to_process:
                frozen.append((list(get_module_recursive(model, path)), f_outs, f_dim))
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
            # Towers this prune() has already swept are frozen: their last block's output for batch i is what the sweep left
            # behind, so their blocks answer from it instead of being run again (the reference re-runs the whole pruned ViT
            # for the T5 encoder's AND the decoder's capture pass: 2 x 128 eager ViT-g forwards in the batch-size-1 recipe).
            patched = []
            for f_layers, f_outs, f_dim in frozen:
                o = f_outs[i] if i < len(f_outs) else None
                if torch.is_tensor(o) and o.shape[f_dim] == spec.batch_len(batch):
                    for blk in f_layers:
                        blk.forward = (lambda *a, _o=o, **k: _o)
                        patched.append(blk)
            try:
                pruner._call_forward_to_cache(model, batch, device)
            except ValueError:
                pass
            finally:
                for blk in patched:
                    del blk.forward
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


class _BlockReplay:
    """N2 -- the stage-2 block forward (wanda_pruner.py:250-253,281-285,530-533,560-563; sparsegpt_pruner.py:355-358,392-395).

    The reference calls every block 2 x n_batches times from Python; with the LAVIS SparseGPT recipe (batch size 1) that is
    256 eager forwards per block and the GPU idles behind the host.  Here G consecutive calibration batches form a group:
    the G block forwards of a group are captured ONCE per block in a CUDA graph over static input / kwarg / output buffers
    and every group of both passes is a replay (the weights are pruned in place between the passes, so the same graph
    serves the second pass).  The kernels of a sample are the ones the eager call launches on the same data: the block
    outputs are bit-identical to the eager sweep.  While the graph is captured the Linear hooks copy their input into a
    staging buffer per DISTINCT input tensor (q / k / v hooks see one tensor and share one buffer, so the accumulators' own
    de-duplication keeps working); after every replay of the first pass the group's staged inputs are copied to a buffer
    that holds the hook inputs of ALL samples (the eager sweep keeps them alive as well, for its deferred launches), and at
    the end of the pass ``add_batch`` is called per sample and Linear in the eager order -- the accumulators see the same
    calls on the same data, so norms and Hessians are bit-identical too.  Anything irregular -- ragged shapes, per-sample kwargs that are not
    tensors of one shape, a block adapter (CLIP), a forward that cannot be captured -- keeps the eager path.
    ECF_BLOCK_GRAPH=0 switches it off, ECF_BLOCK_GRAPH_GROUP sets G (default 16)."""

    last_stats = None
    _keep = None  # (memory pool handle, the most recent graph: keeps the pool alive between blocks)

    def __init__(self, layer, spec, autocast, subset, permute_names, group, n_groups):
        self.layer, self.spec, self.autocast, self.subset = layer, spec, autocast, subset
        self.permute_names = permute_names
        self.G, self.n_groups = group, n_groups
        self.full = []       # per distinct hook input: [n_groups * G, *x.shape], the first pass' inputs of every sample
        self.filled = 0
        self.graph = None
        self.sin = self.sout = None
        self.scache = {}     # kwarg name -> [G, ...] staging of per-sample tensors
        self.shared = {}     # kwarg name -> the value every sample shares
        self.stage = []      # per distinct hook input: [G, *x.shape]
        self.stage_of = {}   # Linear name -> index into stage
        self.failed = False
        self.stats = {"captures": 0, "replays": 0}
        _BlockReplay.last_stats = self.stats

    @staticmethod
    def eligible(inps, caches, batches, spec, group):
        # a graph costs about as much host time to capture as its G forwards cost eagerly: it pays off from the fourth
        # replay on (two groups x two passes) -- the batch-size-1 recipes; 16 batches of 8 stay eager
        if os.environ.get("ECF_BLOCK_GRAPH", "1") == "0" or spec.block_adapter is not None or group < 2 or len(batches) < 2 * group \
                or len(batches) % group != 0:
            return False
        x0 = inps[batches[0]]
        if not (torch.is_tensor(x0) and x0.is_cuda):
            return False
        for j in batches:
            x = inps[j]
            if not torch.is_tensor(x) or x.shape != x0.shape or x.dtype != x0.dtype or x.device != x0.device:
                return False
            if caches[j].keys() != caches[batches[0]].keys():
                return False
        for k, v0 in caches[batches[0]].items():
            for j in batches:
                v = caches[j][k]
                if torch.is_tensor(v0):
                    if not torch.is_tensor(v) or v.shape != v0.shape or v.dtype != v0.dtype or v.device != v0.device:
                        return False
                elif v is not v0:
                    try:  # (a tuple of tensors, e.g. rotary position embeddings, has no truth value: keep the eager path)
                        if not bool(v == v0):
                            return False
                    except Exception:
                        return False
        return True

    def _kwargs(self, g):
        kw = dict(self.shared)
        for k, buf in self.scache.items():
            kw[k] = buf[g]
        return kw

    def _capture(self, inps, caches, chunk):
        G, c0 = self.G, caches[chunk[0]]
        with torch.no_grad(), self.autocast():  # warm-up: lazy initialisations, and the shape of the block's output
            out0 = _run_block(self.layer, inps[chunk[0]], c0, self.spec)
        x0 = inps[chunk[0]]
        self.sin = torch.empty((G,) + tuple(x0.shape), dtype=x0.dtype, device=x0.device)
        self.sout = torch.empty((G,) + tuple(out0.shape), dtype=out0.dtype, device=out0.device)
        del out0
        for k, v in c0.items():
            # a tensor that every sample shares (same object) stays a plain argument; per-sample tensors are staged
            if torch.is_tensor(v) and any(caches[j][k] is not v for j in chunk):
                self.scache[k] = torch.empty((G,) + tuple(v.shape), dtype=v.dtype, device=v.device)
            else:
                self.shared[k] = v
        slot = {"g": 0, "seen": {}}

        def make_stage_hook(name):
            permute = name in self.permute_names

            def hook(_, inp, out):
                x = inp[0].detach()
                if permute:
                    x = x.permute(1, 0, 2)
                key = (x.data_ptr(), tuple(x.shape), tuple(x.stride()), x.dtype)
                # (the tensor itself is kept until the sample's forward ends: a freed activation's address may be handed
                # to a later, different activation of the same shape, which would then pass for a shared input)
                idx = slot["seen"].get(key, (None, None))[0]
                if idx is None:
                    if slot["g"] == 0:
                        idx = len(self.stage)
                        self.stage.append(torch.empty((G,) + tuple(x.shape), dtype=x.dtype, device=x.device))
                        slot.setdefault("order", []).append(name)
                    else:  # the same hook opened this distinct input in slot 0
                        idx = self.stage_of[name]
                    slot["seen"][key] = (idx, x)
                    self.stage[idx][slot["g"]].copy_(x)
                if slot["g"] == 0:
                    self.stage_of[name] = idx
                elif self.stage_of.get(name) != idx:
                    raise RuntimeError("hook inputs are shared differently from sample to sample")
            return hook

        handles = [self.subset[name].register_forward_hook(make_stage_hook(name)) for name in self.subset]
        try:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            # One memory pool for the graphs of all blocks: a private pool per block costs a cudaMalloc / cudaFree round trip
            # of hundreds of MB per block (measured: the BLIP-2 run got slower than the eager sweep).  A pool dies with the last
            # graph that uses it, so the previous block's graph is kept until this one exists.
            if _BlockReplay._keep is None:
                _BlockReplay._keep = (torch.cuda.graph_pool_handle(), None)

            def body():
                with torch.no_grad(), self.autocast():
                    for g in range(G):
                        slot["g"], slot["seen"] = g, {}
                        self.sout[g].copy_(_run_block(self.layer, self.sin[g], self._kwargs(g), self.spec))

            graphs.capture(graph, body, pool=_BlockReplay._keep[0], device=self.sin.device)
        finally:
            for h in handles:
                h.remove()
        self.graph = graph
        _BlockReplay._keep = (_BlockReplay._keep[0], graph)
        self.stats["captures"] += 1

    def run(self, inps, outs, caches, chunk, first):
        """the G block forwards of `chunk`; in the first pass the staged hook inputs of the group are kept"""
        if self.graph is None:
            self._capture(inps, caches, chunk)
        torch.stack([inps[j] for j in chunk], out=self.sin)
        for k, buf in self.scache.items():
            torch.stack([caches[j][k] for j in chunk], out=buf)
        self.graph.replay()
        self.stats["replays"] += 1
        for g, j in enumerate(chunk):
            outs[j] = self.sout[g].clone()
        if first:
            # keep the group's hook inputs: the accumulators get every sample at the end of the pass, exactly as the eager
            # sweep hands them over (NormBatch / HessianBatch defer their launches to one flush per block anyway)
            c = self.filled
            for idx, st in enumerate(self.stage):
                if len(self.full) <= idx:
                    self.full.append(torch.empty((self.n_groups * self.G,) + tuple(st.shape[1:]), dtype=st.dtype, device=st.device))
                self.full[idx][c * self.G:(c + 1) * self.G].copy_(st)
            self.filled += 1

    def finish(self, wrapped):
        """end of the first pass: the very add_batch calls of the eager hooks, in the same order, on the kept inputs"""
        assert self.filled == self.n_groups
        norm = {name: acc for name, acc in wrapped.items() if isinstance(acc, WrappedGPT)}
        for s_ in range(self.n_groups * self.G if norm else 0):
            for name, acc in norm.items():
                acc.add_batch(self.full[self.stage_of[name]][s_], None)
        for name, acc in wrapped.items():
            if name not in norm:
                # Hessians: HessianBatch concatenates the per-sample calls into one launch over [sum T, C]; the kept buffer IS
                # that concatenation ([N, b, L, C] -> [N * b, L, C]; a [N, L, C] stack counts one sample per 2-D input)
                full = self.full[self.stage_of[name]]
                acc.add_batch(full.flatten(0, 1) if full.dim() == 4 else full, None)


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

        group = max(2, int(os.environ.get("ECF_BLOCK_GRAPH_GROUP", "16")))
        while group > 2 and len(my_batches) % group != 0:  # the largest group size <= the setting that divides the batch count
            group -= 1
        replay = None
        if not getattr(pruner, "_block_graph_failed", False) and _BlockReplay.eligible(inps, caches, my_batches, spec, group):
            replay = _BlockReplay(layer, spec, autocast, subset, set(), group, len(my_batches) // group)
        chunks = [my_batches[c:c + group] for c in range(0, len(my_batches), group)]

        def flush():
            if norm_batch is not None:  # one launch for the deferred hook calls
                norm_batch.flush()
            if hess_batch is not None:
                hess_batch.flush()

        def forward_pass(first):
            nonlocal replay
            for ci, chunk in enumerate(chunks):
                if replay is not None:
                    try:
                        replay.run(inps, outs, caches, chunk, first)
                        continue
                    except Exception as exc:
                        if not (first and ci == 0 and replay.stats["replays"] == 0):
                            raise  # outputs of earlier groups are already in place: no clean way back
                        print(f"[ecoflap_b200] block forward not captured in a CUDA graph ({type(exc).__name__}: {exc}); running eagerly")
                        torch.cuda.synchronize()
                        pruner._block_graph_failed = True
                        replay = None
                handles = [subset[name].register_forward_hook(make_hook(name)) for name in wrapped] if first else []
                try:
                    for j in chunk:
                        with torch.no_grad():
                            with autocast():
                                outs[j] = _run_block(layer, inps[j], caches[j], spec)
                finally:
                    for h in handles:
                        h.remove()
            if first:
                if replay is not None:
                    replay.finish(wrapped)
                flush()

        try:
            forward_pass(True)
            if world > 1:  # global running means over the batches of all ranks (one all-reduce for the block)
                if method == "wanda":
                    edist.sync_block_norms(list(wrapped.values()))
                else:
                    edist.sync_block_hessians(list(wrapped.values()))
        finally:
            pass

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

        forward_pass(False)
        del replay
        inps, outs = outs, inps

    if stem is not None:
        stem.config.use_cache = use_cache
    # multi-tower pruners (BLIP-2) keep the swept tower's final outputs for the capture passes of the towers that follow
    # (capture_block_inputs): tensor-returning blocks, every batch on this rank
    if getattr(pruner, "_frozen_towers", None) is not None and world == 1 and spec.block_output_index is None \
            and spec.block_adapter is None and all(torch.is_tensor(t) for t in inps[:n_batches]):
        pruner._frozen_towers.append((module_to_process, inps[:n_batches], spec.sample_dim))
    _BlockReplay._keep = None  # the graphs' memory pool goes back to the allocator
    # (the reference ends _prune with torch.cuda.empty_cache(), wanda_pruner.py:289: with GBs of cached blocks that is 0.6 s of
    # cudaFree per tower and the next tower pays the cudaMalloc again; ECF_EMPTY_CACHE=1 restores it)
    if torch.cuda.is_available() and os.environ.get("ECF_EMPTY_CACHE", "0") == "1":
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
