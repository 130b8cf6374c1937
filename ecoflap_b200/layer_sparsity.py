"""ECoFLaP stage 1: global importance (zeroth-/first-order) and the per-group sparsity allocation.

Mirror of ``LayerSparsity`` (LAVIS/lavis/compression/pruners/layer_single_base_pruner.py:120-561; the CoOp and
UPop copies are identical apart from imports) with the same constructor, ``return_sparsity()``,
``compute_the_sparsity_per_group()`` and ``zo_perturb_parameters()`` seams.

What runs where
  * the perturbation  w += a*eps*z  is the ``ecf_zo_perturb`` kernel (three roundings reproduced, z drawn by
    torch after ``torch.manual_seed`` so the RNG stream is the reference's own);
  * the |W| / W^2 factors of the *Mag* scores are one segmented-reduction launch (``ecf_group_abs_reduce``):
    sum(|W|*g) == g*sum|W| and sum(W^2 g^2) == g^2 * sum W^2, so per-element score tensors are never built;
  * the allocator is host arithmetic on <= a few hundred groups and is restated with torch CPU tensors so
    that the reference's implicit dtype promotions (int64 * float -> fp32, int64 + fp32 -> fp32) and its
    '+=' overshoot quirk (:301) are reproduced bit for bit.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


class _UniformSparsity:
    """``uniform_sparsity_module`` (layer_single_base_pruner.py:327-331): any key -> the global ratio."""

    def __init__(self, ratio):
        self.ratio = ratio

    def __getitem__(self, key):
        return self.ratio


def _map_struct(x, fn):
    if torch.is_tensor(x):
        return fn(x)
    if isinstance(x, (list, tuple)):
        return type(x)(_map_struct(v, fn) for v in x)
    if isinstance(x, dict):
        return {k: _map_struct(v, fn) for k, v in x.items()}
    return x


def _copy_struct(dst, src):
    if torch.is_tensor(dst):
        dst.copy_(src)
    elif isinstance(dst, (list, tuple)):
        for d, s_ in zip(dst, src):
            _copy_struct(d, s_)
    elif isinstance(dst, dict):
        for k in dst:
            _copy_struct(dst[k], src[k])


class _ReplayedLoss:
    """The zeroth-order loop evaluates the SAME no-grad forward 2 * #layers times per first-stage batch (588 layers x 4
    batches x 2 for BLIP-2); in eager mode each of those forwards is ~2 000 kernel launches issued by Python, and the GPU
    idles behind the host.  The first evaluation of a batch is therefore captured in a CUDA graph (after a warm-up run
    on a side stream) and every later one is a replay: the same kernels on the same -- in place perturbed -- parameters,
    hence the same losses, without the launch overhead.  The batch is moved to the device once, before the capture (a
    pageable host-to-device copy cannot be captured).  Anything that cannot be captured (a model that synchronises or
    allocates on the host inside its forward) makes this wrapper fall back to calling ``loss_func`` directly for the rest
    of the run.  ECF_ZO_GRAPH=0 switches it off.

    **Prefix cache** (ECF_ZO_PREFIX=0 switches it off).  Perturbing a layer of block b leaves the outputs of every block
    that runs before b untouched, yet the plain loop recomputes them 2 * #layers times -- for BLIP-2 the 39 ViT-g blocks
    (2 056 tokens per batch) are ~80 % of a forward and 336 of the 588 layers sit behind them in the T5.  The blocks
    (children of an ``nn.ModuleList`` that hold a scored parameter) are put in execution order by the warm-up run; a graph
    VARIANT with cut c returns the cached output of the blocks before c (their ``forward`` is replaced while the variant is
    captured) and recomputes the rest, and every variant copies the outputs of the boundary blocks (c - 1 for every cut
    c, and the last block of every ModuleList) into the cache as part of the graph.  ``dirty_from`` tracks, per batch, the
    first block whose cached output does not correspond to the current weights: after an evaluation for a layer of block b
    the blocks before b were recomputed with clean weights (their caches are fresh) and everything from b on is dirty, so
    the next evaluation may cut at the largest cut <= min(b', dirty_from).  The cached tensors are what the full forward
    would compute at that moment (same kernels, same inputs, including the rounding residue the earlier layers carry after
    their +1 / -2 / +1 cycle), so losses and ratios are bit-identical to the uncached loop."""

    last = None  # the most recent instance (tests read its statistics)

    def __init__(self, loss_func, model, device, names=None):
        import os

        self.loss_func, self.model, self.device = loss_func, model, device
        self.enabled = (os.environ.get("ECF_ZO_GRAPH", "1") != "0" and torch.cuda.is_available()
                        and torch.device(device).type == "cuda")
        self._graphs = {}
        self._batches = {}      # id(batch) -> (device batch, host batch kept alive)
        self._pool = None
        # prefix cache
        self.prefix = self.enabled and os.environ.get("ECF_ZO_PREFIX", "1") != "0" and names is not None
        self._stride = max(1, int(os.environ.get("ECF_ZO_PREFIX_STRIDE", "4")))
        self._blocks = None     # block modules in execution order
        self._block_of = {}     # parameter name -> index into _blocks
        self._cuts = [0]
        self._boundary = set()  # block indices whose output is cached
        self._ret = {}          # cut -> {skipped block index: boundary block index whose cache it returns}
        self._cache = {}        # id(batch) -> {boundary block index: cached output}
        self._dirty = {}        # id(batch) -> first block whose cache is not valid for the current weights
        self._own = set()       # blocks that keep their own (never refreshed) output as a structural stand-in
        # static buffers: one set of graph variants for all batches of one shape (ECF_ZO_STATIC=0: one set per batch)
        self.static = self.enabled and os.environ.get("ECF_ZO_STATIC", "1") != "0"
        self._sbatch, self._scache = None, {}
        self._static_keys = set()
        self._cur, self._newer, self._loaded = None, set(), set()
        self.stats = {"replays": 0, "cuts_used": set(), "captures": 0}
        _ReplayedLoss.last = self  # (tests look at the statistics of the most recent run)
        if self.prefix:
            self._find_blocks(names)

    # -- prefix cache: structure ------------------------------------------------------------------------------------
    def _find_blocks(self, names):
        import torch.nn as nn

        mods = dict(self.model.named_modules())
        cand, owner = [], {}
        for name in names:
            parts = name.split(".")
            blk = None
            for j in range(1, len(parts)):  # OUTERMOST ancestor that is a child of a ModuleList (a T5Block, not its sub-layers)
                parent = mods.get(".".join(parts[:j - 1])) if j > 1 else self.model
                me = mods.get(".".join(parts[:j]))
                if isinstance(parent, nn.ModuleList) and me is not None:
                    blk = me
                    break
            if blk is not None:
                owner[name] = blk
                if all(blk is not c for c in cand):
                    cand.append(blk)
        if len(cand) < 2:
            self.prefix = False
            return
        self._cand, self._owner = cand, owner

    def _order_blocks(self, dev_batch):
        """One eager forward with pre-hooks on the candidate blocks: execution order; every block must run exactly once."""
        import torch.nn as nn

        order, handles = [], []
        for blk in self._cand:
            handles.append(blk.register_forward_pre_hook(lambda m, a, _o=order: _o.append(m)))
        try:
            with torch.no_grad():
                self.loss_func(self.model, dev_batch, True)
        finally:
            for h in handles:
                h.remove()
        if len(order) != len(self._cand) or len({id(m) for m in order}) != len(order):
            self.prefix = False
            return
        self._blocks = order
        index = {id(m): i for i, m in enumerate(order)}
        self._block_of = {n: index[id(b)] for n, b in self._owner.items()}
        # towers: maximal runs of consecutive blocks that share their parent ModuleList
        parent_of = {}
        for mod in self.model.modules():
            if isinstance(mod, nn.ModuleList):
                for ch in mod:
                    parent_of[id(ch)] = id(mod)
        towers, start = [], 0
        for i in range(1, len(order) + 1):
            if i == len(order) or parent_of.get(id(order[i])) != parent_of.get(id(order[start])):
                towers.append((start, i - 1))
                start = i
        cuts = {0}
        for t0, t1 in towers:
            cuts.update(range(t0, t1 + 1, self._stride))
        self._cuts = sorted(cuts)
        self._boundary = {c - 1 for c in self._cuts if c > 0} | {t1 for _, t1 in towers}
        for c in self._cuts:
            ret = {}
            for t0, t1 in towers:
                last = t1 if t1 < c else (c - 1 if t0 < c else None)
                if last is None:
                    continue
                # a skipped block answers with the boundary block's cache when both are the same kind of module (same output
                # structure; its value only feeds the next skipped block), else with its own first output
                ret.update({j: (last if type(order[j]) is type(order[last]) else j) for j in range(t0, last + 1)})
            self._ret[c] = ret
        self._own = {j for ret in self._ret.values() for j, src in ret.items() if src == j} - self._boundary

    # -- running one variant ----------------------------------------------------------------------------------------
    def _run(self, cache, cut, dev_batch):
        """loss_func with the blocks before `cut` answering from `cache` and the boundary blocks from `cut` on writing to it"""
        patched, handles = [], []
        try:
            if self.prefix and self._blocks is not None:
                for j, src in self._ret[cut].items():
                    blk = self._blocks[j]
                    blk.forward = (lambda *a, _c=cache[src], **k: _c)
                    patched.append(blk)
                for j in self._boundary | self._own:
                    if j in self._ret[cut]:
                        continue

                    def hook(mod, inp, out, _j=j, _update=j in self._boundary):
                        if _j not in cache:
                            cache[_j] = _map_struct(out, lambda t: t.detach().clone())
                        elif _update:
                            _copy_struct(cache[_j], out)
                    handles.append(self._blocks[j].register_forward_hook(hook))
            return self.loss_func(self.model, dev_batch, True)
        finally:
            for blk in patched:
                del blk.forward
            for h in handles:
                h.remove()

    @staticmethod
    def _same_struct(a, b):
        """same nesting, tensors of the same shape / dtype, everything else equal"""
        if torch.is_tensor(a) or torch.is_tensor(b):
            return torch.is_tensor(a) and torch.is_tensor(b) and a.shape == b.shape and a.dtype == b.dtype and a.device == b.device
        if isinstance(a, (list, tuple)):
            return type(a) is type(b) and len(a) == len(b) and all(_ReplayedLoss._same_struct(x, y) for x, y in zip(a, b))
        if isinstance(a, dict):
            return isinstance(b, dict) and a.keys() == b.keys() and all(_ReplayedLoss._same_struct(a[k], b[k]) for k in a)
        try:
            return bool(a == b)
        except Exception:
            return False

    @staticmethod
    def _to_device(x, device):
        if torch.is_tensor(x):
            return x.to(device)
        if isinstance(x, dict):
            return {k: _ReplayedLoss._to_device(v, device) for k, v in x.items()}
        if isinstance(x, (list, tuple)):
            return type(x)(_ReplayedLoss._to_device(v, device) for v in x)
        return x

    def _eager(self, batch):
        with torch.no_grad():
            return self.loss_func(self.model, batch, self.device != "cpu")

    def __call__(self, batch, name=None):
        """-> (loss as a fresh 0-d tensor, batch_len).  ``name``: the parameter that is perturbed right now (prefix cache)."""
        if not self.enabled:
            return self._eager(batch)
        key = id(batch)
        if key not in self._batches and len(self._batches) >= 64:
            # a loader that builds new batch objects on every pass never hits the cache: stop capturing
            print("[ecoflap_b200] zeroth-order loader yields fresh batch objects on every pass; running eagerly")
            self.enabled = False
            self._graphs.clear()
            return self._eager(batch)
        try:
            if key not in self._batches:
                dev_batch = self._to_device(batch, self.device)
                side = torch.cuda.Stream(device=self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side), torch.no_grad():
                    if self.prefix and self._blocks is None:
                        self._order_blocks(dev_batch)
                    self._run(self._cache.setdefault(key, {}), 0, dev_batch)  # warm-up; fills the caches of a new batch
                torch.cuda.current_stream(self.device).wait_stream(side)
                torch.cuda.synchronize(self.device)
                self._batches[key] = (dev_batch, batch)  # (the host batch is kept alive: ids are the keys)
                self._dirty[key] = 0
                if self.static and self._sbatch is None:  # the first batch shapes the static buffers
                    self._sbatch = _map_struct(dev_batch, lambda t: t.clone())
                    self._scache = {j: _map_struct(v, lambda t: t.clone()) for j, v in self._cache[key].items()}
                    self._cur, self._newer = key, set()
                    self._loaded = set(self._scache)
                if self.static and self._same_struct(self._sbatch, dev_batch) and \
                        all(self._same_struct(self._scache.get(j), v) for j, v in self._cache[key].items()):
                    self._static_keys.add(key)
            dev_batch = self._batches[key][0]
            cut = 0
            blk = self._block_of.get(name) if (self.prefix and self._blocks is not None) else None
            if blk is not None:
                lim = min(blk, self._dirty[key])
                cut = max(c for c in self._cuts if c <= lim)
            static = key in self._static_keys
            if static:
                # ONE set of graph variants over static batch / cache buffers serves every batch of the same shape: a
                # capture (the suffix forward in capture mode plus the instantiation of ~1 000 nodes) costs 0.15-0.3 s,
                # and per-batch variants made 88 of them per run.  Switching batches copies the batch in, the boundary
                # outputs the previous batch's replays refreshed back out, and the ones this variant reads in.
                if self._cur != key:
                    old = self._cache[self._cur]
                    for j in self._newer:
                        _copy_struct(old[j], self._scache[j])
                    _copy_struct(self._sbatch, dev_batch)
                    self._cur, self._newer, self._loaded = key, set(), set()
                if self.prefix and self._blocks is not None:
                    for src in set(self._ret[cut].values()):
                        if src not in self._loaded and src not in self._newer:
                            _copy_struct(self._scache[src], self._cache[key][src])
                            self._loaded.add(src)
            gkey = ("static", cut) if static else (key, cut)
            entry = self._graphs.get(gkey)
            if entry is None:
                if self._pool is None:
                    self._pool = torch.cuda.graph_pool_handle()
                torch.cuda.synchronize(self.device)
                graph = torch.cuda.CUDAGraph()
                # (torch's context manager, allocator flush included: measured 47.2 s against 50.1 s with the flush-free
                # capture of graphs.py on the same box -- the other way round from the stage-2 sweep)
                with torch.cuda.graph(graph, pool=self._pool), torch.no_grad():
                    if static:
                        loss, batch_len = self._run(self._scache, cut, self._sbatch)
                    else:
                        loss, batch_len = self._run(self._cache[key], cut, dev_batch)
                entry = (graph, loss, int(batch_len))
                self._graphs[gkey] = entry
                self.stats["captures"] += 1
        except Exception as exc:  # not capturable: say so once, continue eagerly
            print(f"[ecoflap_b200] zeroth-order forward not captured in a CUDA graph ({type(exc).__name__}: {exc}); running eagerly")
            torch.cuda.synchronize(self.device)
            self.enabled = False
            self._graphs.clear()
            return self._eager(batch)
        graph, loss, batch_len = entry
        graph.replay()
        if static and self.prefix and self._blocks is not None:  # boundary blocks this variant recomputed: newer than the batch's copy
            upd = {j for j in self._boundary if j not in self._ret[cut]}
            self._newer |= upd
            self._loaded -= upd
        self.stats["replays"] += 1
        self.stats["cuts_used"].add(cut)
        # blocks in [cut, blk) were recomputed with the current weights: their caches are fresh; from blk on they are not
        self._dirty[key] = blk if blk is not None else 0
        return loss.clone(), batch_len


class LayerSparsity:
    # Device the zeroth-order noise is drawn on.  None = the parameter's own device, which is what the reference does
    # (layer_single_base_pruner.py:479: CUDA Philox stream in a GPU run, mt19937 on CPU).  Setting "cpu" draws the
    # reference's CPU stream and moves z to the parameter -- used to check the loop against CPU-generated fixtures.
    noise_device = None

    def __init__(self, model, data_loader, loss_func, num_samples, original_sparsity, max_sparsity_per_layer=0.8,
                 score_method="GradMagSquare_avg", num_noise=1, noise_eps=1e-3, layer_to_group_mapping={},
                 prune_per_model=False, per_model_group=[]):
        self.importance_measure = {}
        self.model = model
        self.data_loader = data_loader
        self.loss_func = loss_func
        self.num_samples = num_samples
        self.original_sparsity = original_sparsity
        self.layer_to_group_mapping = layer_to_group_mapping
        self.max_sparsity_per_layer = max_sparsity_per_layer
        self.num_noise = num_noise
        self.noise_eps = noise_eps
        self.prune_per_model = prune_per_model
        self.score_method = score_method
        self.per_model_group = per_model_group
        if score_method is not None:
            self.score_compute, self.score_aggregate = score_method.split("_")
        assert self.max_sparsity_per_layer >= self.original_sparsity

    # ------------------------------------------------------------------ A15 allocator (host)
    def compute_the_sparsity_per_group(self, total_parameters_to_keep, group_scores, group_num_parameters,
                                       max_sparsity_per_layer=0.8):
        """Water-filling allocation, arithmetic identical to layer_single_base_pruner.py:247-314."""
        target = total_parameters_to_keep
        score = torch.FloatTensor(list(group_scores.values()))
        size = torch.LongTensor(list(group_num_parameters.values()))
        keep_frac = 1 - max_sparsity_per_layer

        # floor guaranteeing the per-group maximum sparsity: ceil(int64 * float) is evaluated in fp32
        keep = torch.zeros_like(score, dtype=int)
        keep += torch.ceil(size * keep_frac).int()

        while keep.sum() < target:
            missing = target - keep.sum()
            grant = torch.ceil((score / torch.sum(score)) * missing)
            keep = keep + grant  # int64 + fp32 -> fp32 from the first round on
            score[keep >= size] = 0  # saturated groups stop competing
            keep = torch.clamp(keep, max=size)

            if grant.sum() == 0:  # nothing could be granted: hand the shortfall out in index order
                have = keep.sum()
                if have < target:
                    short = target - have
                    while short > 0:
                        for g in torch.where(score > 0)[0]:
                            room = min(short, size[g] - keep[g])
                            keep[g] += room
                            short -= room
                            if short == 0:
                                break

            if keep.sum() > target:  # overshoot: the reference ADDS the removable amount (sic)
                excess = keep.sum() - target
                while excess > 0:
                    for g in torch.argsort(keep, descending=True, stable=True):
                        removable = min(excess, keep[g] - (size[g] * keep_frac).int())
                        keep[g] += removable
                        excess -= removable
                        if excess == 0:
                            break

        return {
            name: torch.clamp(1 - k / n, min=0, max=1).item()
            for name, k, n in zip(group_num_parameters.keys(), keep, size)
        }

    # ------------------------------------------------------------------ A14 + A15 driver
    def return_sparsity(self):
        mapping = self.layer_to_group_mapping
        if self.score_compute.startswith("Real"):
            # the 3-iteration global pruning run as a ratio oracle (layer_single_base_pruner.py:321-325)
            return self.global_iterative_pruning(self.original_sparsity, mapping, iteratation=3, max_sparsity_per_layer=1.0)
        if mapping is None or len(mapping) == 0:
            return _UniformSparsity(self.original_sparsity)

        if len(self.importance_measure) == 0:
            if self.score_compute.startswith("MEZO"):
                self.importance_measure = self.compute_importance_scores_mezo(mapping)
            else:
                self.importance_measure = self.compute_importance_scores(mapping)

        numel = {k: v.numel() for k, v in self.model.named_parameters() if k in mapping}
        total = sum(numel.values())
        total_to_keep = int(total * (1 - self.original_sparsity))

        group_scores, group_sizes = {}, {}
        for layer, group in mapping.items():
            if group not in group_scores:
                group_scores[group] = 0
                group_sizes[group] = 0
        members = {}
        for layer, group in mapping.items():
            members.setdefault(group, []).append(layer)
        for group, layers in members.items():
            for layer in layers:
                group_scores[group] += self.importance_measure[layer].sum()
                group_sizes[group] += numel[layer]
            if self.score_aggregate == "avg":
                group_scores[group] /= group_sizes[group]

        if self.prune_per_model:
            group_sparsity = {}
            for prefix in self.per_model_group:
                sub_scores = {k: v for k, v in group_scores.items() if k.startswith(prefix)}
                sub_sizes = {k: v for k, v in group_sizes.items() if k.startswith(prefix)}
                sub_keep = int(sum(list(sub_sizes.values())) * (1 - self.original_sparsity))
                group_sparsity.update(self.compute_the_sparsity_per_group(
                    sub_keep, sub_scores, sub_sizes, max_sparsity_per_layer=self.max_sparsity_per_layer))
        else:
            group_sparsity = self.compute_the_sparsity_per_group(
                total_to_keep, group_scores, group_sizes, max_sparsity_per_layer=self.max_sparsity_per_layer)

        kept = sum((1 - group_sparsity[g]) * group_sizes[g] for g in group_sizes)
        print(kept, total_to_keep)  # the reference's sanity line (:407)
        return {layer: group_sparsity[group] for layer, group in mapping.items()}

    # ------------------------------------------------------------------ 'Real*' ratio oracle (N3 machinery)
    def global_iterative_pruning(self, target_sparsity, dict_layers_to_prune, iteratation=1, max_sparsity_per_layer=1.0):
        """layer_single_base_pruner.py:183-245: prune globally by the first-order score in ``iteratation`` rounds
        (p_i = target ** (iteratation / i)), read every parameter's zero fraction, restore the weights.  The per-element
        scores are never built: the gradient sums stay on the device and the select recomputes |w| * |g| on the fly
        (pruners/global_pruner.py, csrc/global_select.cu).  Reference detail kept: the gradient accumulation squares only
        when score_compute == "GradMagSquare" EXACTLY (:447), which a 'Real*' name never is, so 'RealGradMagSquare' scores
        w^2 * mean|g|."""
        from .pruners.global_pruner import accumulate_abs_grads, device_get_mask_

        names, params = self._selected(dict_layers_to_prune)
        weight_copy = [p.data.clone() for p in params]
        if "GradMagSquare" in self.score_compute:
            mode = "grad_mag_sq"
        elif "GradMagAbs" in self.score_compute:
            mode = "grad_mag_abs"
        elif "GradOnly" in self.score_compute:
            mode = "grad_only"
        else:
            raise ValueError(f"unknown first-order score method {self.score_compute!r}")
        for i in range(1, iteratation + 1):
            p_i = target_sparsity ** (iteratation / i)
            G, nb = accumulate_abs_grads(self.model, self.data_loader, self.loss_func, names, params, self.num_samples,
                                         square=(self.score_compute == "GradMagSquare"))
            device_get_mask_(params, G, nb, mode, p_i, max_sparsity_per_layer)
            del G
            print(f"Step {i}, target sparsity: {p_i:.4f}")
        sparsity_dict = {}
        for k, v in self.model.named_parameters():
            z = int(ops.count_zero(v.data if v.data.is_contiguous() else v.data.contiguous()).item())
            sparsity_dict[k] = (torch.tensor(float(z), dtype=torch.float32) / v.numel()).item()
        for p, w in zip(params, weight_copy):
            p.data = w
        return sparsity_dict

    # ------------------------------------------------------------------ helpers
    def _selected(self, mapping):
        names, params = [], []
        for k, v in self.model.named_parameters():
            if k in mapping:
                names.append(k)
                params.append(v)
        return names, params

    def _grad_accumulate(self, G, grads, square):
        """G += |g| (or g^2) per parameter, fp32, on the device: ecf_grad_accum."""
        for g_acc, gr in zip(G, grads):
            ops.grad_accum(g_acc, gr.detach(), square=square)

    def _score_sums(self, params, G, n_batches, mode):
        """sum over the elements of the first-order score per parameter as a float64 vector: ecf_global_score_sum."""
        datas = [p.data if p.data.is_contiguous() else p.data.contiguous() for p in params]
        return ops.GlobalTable(datas, G, n_batches, mode).score_sums()

    def _magnitude_sums(self, params):
        """(sum|w|, sum w^2) per parameter as python floats: one kernel launch over all tensors."""
        sa, sq = ops.group_abs_reduce([p.data for p in params])
        return sa.cpu().tolist(), sq.cpu().tolist()

    # ------------------------------------------------------------------ A11 perturbation
    def zo_perturb_parameters(self, params, random_seed=1, scaling_factor=1, zo_eps=1e-3):
        """theta <- theta + scaling_factor * z * zo_eps with z ~ N(0, 1) drawn after torch.manual_seed
        (layer_single_base_pruner.py:473-486); the update itself is the ecf_zo_perturb kernel."""
        torch.manual_seed(random_seed)
        for param in params:
            if self.noise_device is None:
                z = torch.normal(mean=0, std=1, size=param.data.size(), device=param.data.device, dtype=param.data.dtype)
            else:
                z = torch.normal(mean=0, std=1, size=param.data.size(), device=self.noise_device,
                                 dtype=param.data.dtype).to(param.data.device)
            data = param.data if param.data.is_contiguous() else param.data.contiguous()
            ops.zo_perturb(data, z, scaling_factor, zo_eps)
            if data is not param.data:
                param.data.copy_(data)

    # ------------------------------------------------------------------ A12 zeroth-order scores
    def compute_importance_scores_mezo(self, layer_to_group_mapping):
        from . import dist as edist

        if edist.is_dist():
            return self._compute_importance_scores_mezo_sharded(layer_to_group_mapping)
        model, loss_func = self.model, self.loss_func
        model.eval()
        names, params = self._selected(layer_to_group_mapping)
        device = next(iter(model.parameters())).device
        eps = self.noise_eps
        ghat = {k: 0.0 for k in names}
        evaluate = _ReplayedLoss(loss_func, model, device, names)
        for i, (name, param) in enumerate(zip(names, params)):
            print(i, name)
            seen = 0
            for batch in self.data_loader:
                if seen >= self.num_samples:
                    break
                acc = 0.0
                for _ in range(self.num_noise):
                    if seen >= self.num_samples:
                        break
                    seed = np.random.randint(1000000000)
                    self.zo_perturb_parameters([param], random_seed=seed, scaling_factor=1, zo_eps=eps)
                    loss_plus, batch_len = evaluate(batch, name)
                    self.zo_perturb_parameters([param], random_seed=seed, scaling_factor=-2, zo_eps=eps)
                    loss_minus, batch_len = evaluate(batch, name)
                    # restore (inexact in fp16/bf16 exactly as in the reference)
                    self.zo_perturb_parameters([param], random_seed=seed, scaling_factor=1, zo_eps=eps)
                    seen += batch_len
                    acc += abs(((loss_plus - loss_minus) / (2 * eps)).item())
                    torch.manual_seed(seed)
                ghat[name] += acc
        return self._mezo_scores(names, params, ghat)

    def _compute_importance_scores_mezo_sharded(self, layer_to_group_mapping):
        """The zeroth-order loop on P ranks (the reference is single-GPU): the (layer, batch) grid is
        embarrassingly parallel over layers, so rank r evaluates the layers l = r (mod P) on a replicated model and one
        all-reduce of the [#layers] g-hat vector ends the stage -- 2 * #layers * #batches forwards become
        2 * #layers * #batches / P per GPU.  What is reproduced of the single-GPU run:
          * the numpy seed stream, pre-drawn in the reference's (layer, batch, noise) order on every rank (the number of
            draws per layer comes from one unperturbed pass over the first-stage batches), hence every z;
          * the weights stage 2 scores: every owner broadcasts its perturbed-and-inexactly-restored parameters afterwards
            (one +1/-2/+1 cycle changes ~45 % of the elements), bit-identical to the sequential loop;
          * g-hat up to the rounding residue of the OTHER ranks' layers: the sequential loop evaluates layer i on a
            model whose layers < i already carry that residue (<= 5e-3 relative in bf16, 7e-8 in fp32), here only the
            rank's own earlier layers do -- both losses of a pair see the same residue, so the effect is second order."""
        import torch.distributed as dist

        from . import dist as edist

        model, loss_func = self.model, self.loss_func
        model.eval()
        names, params = self._selected(layer_to_group_mapping)
        device = next(iter(model.parameters())).device
        eps = self.noise_eps
        rank, world = edist.rank_world()
        # one unperturbed pass: batch lengths -> seed draws per layer
        batch_lens, seen = [], 0
        for batch in self.data_loader:
            if seen >= self.num_samples:
                break
            with torch.no_grad():
                _, batch_len = loss_func(model, batch, device != "cpu")
            batch_lens.append(int(batch_len))
            seen += int(batch_len)
        draws = edist.zo_draws_per_layer(batch_lens, self.num_samples, self.num_noise)
        evaluate = _ReplayedLoss(loss_func, model, device, names)
        seeds = [[int(np.random.randint(1000000000)) for _ in range(draws)] for _ in names]  # the reference's stream
        ghat_vec = torch.zeros(len(names), dtype=torch.float64, device=device)
        mismatch = False
        for i, (name, param) in enumerate(zip(names, params)):
            if edist.zo_owner(i, world) != rank:
                continue
            print(i, name)
            seen, d, total = 0, 0, 0.0
            for batch in self.data_loader:
                if seen >= self.num_samples:
                    break
                acc = 0.0
                for _ in range(self.num_noise):
                    if seen >= self.num_samples:
                        break
                    seed = seeds[i][d]
                    d += 1
                    self.zo_perturb_parameters([param], random_seed=seed, scaling_factor=1, zo_eps=eps)
                    loss_plus, batch_len = evaluate(batch, name)
                    self.zo_perturb_parameters([param], random_seed=seed, scaling_factor=-2, zo_eps=eps)
                    loss_minus, batch_len = evaluate(batch, name)
                    self.zo_perturb_parameters([param], random_seed=seed, scaling_factor=1, zo_eps=eps)
                    seen += batch_len
                    acc += abs(((loss_plus - loss_minus) / (2 * eps)).item())
                total += acc
            if d != draws:
                mismatch = True  # raised on EVERY rank after the collective below, never before it (no hang)
            ghat_vec[i] = total
        flag = torch.tensor([1.0 if mismatch else 0.0], dtype=torch.float64, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if float(flag.item()) != 0.0:
            raise RuntimeError("the first-stage loader yielded different batches on its second pass (shuffling loader?): "
                               "the sharded zeroth-order loop needs a deterministic loader")
        dist.all_reduce(ghat_vec, op=dist.ReduceOp.SUM)
        for i, param in enumerate(params):  # every rank continues with the weights the single-GPU loop would leave
            dist.broadcast(param.data, src=edist.zo_owner(i, world))
        if draws and names:
            torch.manual_seed(seeds[-1][-1])
        ghat = {k: float(ghat_vec[i].item()) for i, k in enumerate(names)}
        return self._mezo_scores(names, params, ghat)

    def _mezo_scores(self, names, params, ghat):
        """layer_single_base_pruner.py:551-559 with the weight factor reduced on the device.  Values are
        1-element fp32 tensors holding sum(score) -- the only thing return_sparsity reads."""
        g = {k: torch.FloatTensor([ghat[k]]).abs() for k in names}
        if self.score_compute == "MEZO-GradOnly":
            return {k: g[k].abs() for k in names}
        sum_abs, sum_sq = self._magnitude_sums(params)
        if self.score_compute == "MEZO-GradMagAbs":
            return {k: torch.FloatTensor([sum_abs[i]]) * g[k].abs() for i, k in enumerate(names)}
        if self.score_compute == "MEZO-GradMagSquare":
            return {k: torch.FloatTensor([sum_sq[i]]) * g[k] ** 2 for i, k in enumerate(names)}
        raise ValueError(f"unknown zeroth-order score method {self.score_compute!r}")

    # ------------------------------------------------------------------ A13 first-order scores
    def compute_importance_scores(self, layer_to_group_mapping):
        """mean over batches of |dL/dW| (or g^2), then |W|*|g| / W^2*g / |g| (:416-471).  Gradients stay on the
        device (the reference copies 3.7 G fp32 elements to the host per batch); only sum(score) is kept."""
        from . import dist as edist

        model, loss_func = self.model, self.loss_func
        names, params = self._selected(layer_to_group_mapping)
        device = next(iter(model.parameters())).device
        square = self.score_compute == "GradMagSquare"
        G = [torch.zeros(p.shape, dtype=torch.float32, device=p.device) for p in params]  # sum of |g| (g^2), on the device
        seen, nbatches = 0, 0
        # P ranks: data parallel over the first-stage batches, rank r takes batch j = r (mod P).  Every
        # score is linear in the per-batch |g| (or g^2) terms, so the ranks exchange one scalar per layer at the end
        # instead of gradients.
        rank, world = edist.rank_world() if edist.is_dist() else (0, 1)
        for j, batch in enumerate(self.data_loader):
            if seen >= self.num_samples:
                break
            if world > 1 and j % world != rank:
                seen += self._batch_len(batch, device)
                nbatches += 1
                continue
            loss, batch_len = loss_func(model, batch, device != "cpu")
            seen += batch_len
            nbatches += 1
            grads = torch.autograd.grad(loss, params)
            assert len(grads) == len(names) == len(params)
            self._grad_accumulate(G, grads, square)
        if "GradMagSquare" in self.score_compute:
            mode = "grad_mag_sq"
        elif "GradMagAbs" in self.score_compute:
            mode = "grad_mag_abs"
        elif "GradOnly" in self.score_compute:
            mode = "grad_only"
        else:
            raise ValueError(f"unknown first-order score method {self.score_compute!r}")
        # sum over the elements of |W| * |G / nb|  (W^2 * G / nb, |G / nb|): one segmented-reduction launch over all the
        # layers (ecf_global_score_sum); the per-element score tensors of the reference (:463-469) are never built
        vec = self._score_sums(params, G, max(1, nbatches), mode)
        del G
        if world > 1:
            import torch.distributed as dist

            dist.all_reduce(vec, op=dist.ReduceOp.SUM)
        vec = vec.float().cpu()
        return {k: vec[i:i + 1].clone() for i, k in enumerate(names)}

    def _batch_len(self, batch, device):
        """Length of a first-stage batch this rank does not evaluate (the stopping rule counts every batch): the
        conventions of the reference's loss closures (utils.py:21-66, CoOp zsclip.py:61-95, UPop tuple batches), else
        one forward."""
        if isinstance(batch, dict):
            for key in ("text_input", "image", "label", "labels", "img", "input_ids"):
                if key in batch:
                    return len(batch[key])
        if isinstance(batch, (list, tuple)) and len(batch) and hasattr(batch[0], "shape"):
            return int(batch[0].shape[0])
        with torch.no_grad():
            return int(self.loss_func(self.model, batch, device != "cpu")[1])
