"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the ECoFLaP pruning hot path.

This file is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The shipped path
(``ecoflap_b200``) never imports anything under ``oracle/`` and has no CPU fallback.

Every function restates one piece of the reference (ylsung/ECoFLaP @ 59dac0a) and cites the
file:line it follows (paths relative to the reference root).  Pinning: the reference ships NO golden
vectors or unit tests for this path, so the oracle is pinned against outputs
of the reference's own classes executed in the build container on seeded synthetic inputs --
``tests/gen_golden.py`` (committed) imports the unmodified reference through
``oracle/ref_loader.py`` and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks
every function below against those fixtures.

Conventions: all arrays are numpy; low-precision tensors travel as float32 arrays holding exactly
representable values plus a dtype tag in {"fp32", "fp16", "bf16"}.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32

# --------------------------------------------------------------------------------------
# dtype helpers
# --------------------------------------------------------------------------------------


def round_bf16(x):
    """Round-to-nearest-even fp32 -> bf16, returned as fp32 (what torch's .to(bfloat16) does)."""
    x = np.ascontiguousarray(x, dtype=F32)
    u = x.view(np.uint32).astype(np.uint64)
    nan = np.isnan(x)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = rounded.view(F32).copy()
    out[nan] = np.nan
    return out.reshape(x.shape)


def round_to(x, dtype):
    if dtype == "fp32":
        return np.asarray(x, dtype=F32)
    if dtype == "fp16":
        with np.errstate(over="ignore"):
            return np.asarray(x, dtype=F32).astype(np.float16).astype(F32)
    if dtype == "bf16":
        return round_bf16(x)
    raise ValueError(dtype)


# --------------------------------------------------------------------------------------
# A1  calibration norm accumulator
# --------------------------------------------------------------------------------------


class NormAccumulator:
    """WrappedGPT.__init__/add_batch -- LAVIS/lavis/compression/pruners/wanda_pruner.py:54-84
    (same code: CoOp/trainers/pruners/wanda_pruner.py:142-172, UPop/pruners/wanda_pruner.py:48-78).

    scaler_row <- scaler_row * n/(n+B);  n <- n+B;  scaler_row <- scaler_row + ||x_c||_2^2 / n
    with B = inp.shape[0] (batch entries, not tokens) and x cast to fp32 first.
    """

    def __init__(self, columns):
        self.columns = int(columns)
        self.scaler_row = np.zeros(self.columns, dtype=F32)
        self.nsamples = 0

    def add_batch(self, inp):
        inp = np.asarray(inp)
        if inp.ndim == 2:
            inp = inp[None]
        b = inp.shape[0]
        x = inp.reshape(-1, inp.shape[-1]).astype(F32)  # [T, C]
        self.scaler_row = self.scaler_row * F32(self.nsamples / (self.nsamples + b))
        self.nsamples += b
        # torch.norm(p=2, dim=1) ** 2 : sqrt of the fp32 sum of squares, then squared again
        nrm = np.sqrt(np.sum(x * x, axis=0, dtype=F32), dtype=F32)
        self.scaler_row = (self.scaler_row + (nrm * nrm) / F32(self.nsamples)).astype(F32)
        return self.scaler_row


def sqnorm_columns(x):
    """Raw per-input-channel sum of squares of a [T, C] block in float64 (tolerance anchor)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1, np.shape(x)[-1])
    return np.sum(x * x, axis=0)


# --------------------------------------------------------------------------------------
# A3-A7  Wanda score + selection + apply
# --------------------------------------------------------------------------------------


def wanda_metric(W, scaler_row):
    """W_metric = |W| * sqrt(scaler_row)[None, :]  -- wanda_pruner.py:260 / :541.  fp32, bit exact:
    one IEEE sqrtf and one fp32 multiply per element."""
    W = np.asarray(W, dtype=F32)
    return (np.abs(W) * np.sqrt(np.asarray(scaler_row, dtype=F32))[None, :]).astype(F32)


def row_k(C, sparsity):
    """int(W_metric.shape[1] * sparsity) -- wanda_pruner.py:276 (python float arithmetic)."""
    return int(C * sparsity)


def row_select_mask(M, k):
    """Per-row variant: stable ascending sort of each row, first k indices are pruned
    (wanda_pruner.py:272-277; CoOp wanda_pruner.py:379-381; UPop wanda_pruner.py:253-258;
    LLaMA/image_classifiers/prune_utils.py:35-38 'row').  Ties -> lower column index."""
    M = np.asarray(M, dtype=F32)
    R, C = M.shape
    mask = np.zeros((R, C), dtype=bool)
    if k <= 0:
        return mask
    idx = np.argsort(M, axis=-1, kind="stable")[:, :k]
    np.put_along_axis(mask, idx, True, axis=-1)
    return mask


def layer_kth_index(numel, sparsity):
    """int(W_metric.numel() * sparsity) -- wanda_pruner.py:555."""
    return int(numel * sparsity)


def layer_thresh_mask(M, kth_index):
    """Per-layer variant: thres = sort(flatten)[kth_index]; mask = M <= thres
    (wanda_pruner.py:553-556; UPop :512-515; prune_utils.py:28-31 'layer').
    Prunes kth_index+1 entries plus all ties; kth_index >= numel raises like the reference."""
    M = np.asarray(M, dtype=F32)
    flat = np.sort(M.reshape(-1), kind="stable")
    if kth_index >= flat.size or kth_index < -flat.size:
        raise IndexError("kth index out of range (reference raises IndexError)")
    thres = flat[kth_index]
    return (M <= thres), thres


def nm_select_mask(M, n, m):
    """n:m structured variant (dead code in every shipped config) -- wanda_pruner.py:265-270.
    torch.topk(largest=False) tie order is unspecified; lowest index first is used here."""
    M = np.asarray(M, dtype=F32)
    R, C = M.shape
    mask = np.zeros((R, C), dtype=bool)
    for ii in range(0, C, m):
        blk = M[:, ii:ii + m]
        idx = np.argsort(blk, axis=-1, kind="stable")[:, :n]
        sub = np.zeros(blk.shape, dtype=bool)
        np.put_along_axis(sub, idx, True, axis=-1)
        mask[:, ii:ii + m] = sub
    return mask


def apply_mask(W, mask):
    """subset[name].weight.data[W_mask] = 0 -- wanda_pruner.py:279 / :558."""
    out = np.array(W, copy=True)
    out[mask] = 0
    return out


def wanda_prune_rows(W, scaler_row, sparsity):
    M = wanda_metric(W, scaler_row)
    mask = row_select_mask(M, row_k(M.shape[1], sparsity))
    return apply_mask(W, mask), mask


def wanda_prune_layer(W, scaler_row, sparsity):
    M = wanda_metric(W, scaler_row)
    mask, thres = layer_thresh_mask(M, layer_kth_index(M.size, sparsity))
    return apply_mask(W, mask), mask, thres


# --------------------------------------------------------------------------------------
# A8  SparseGPT Hessian accumulator
# --------------------------------------------------------------------------------------


class HessianAccumulator:
    """SparseGPT.__init__/add_batch -- sparsegpt_pruner.py:56-82 (CoOp sparsegpt_pruner.py:145-171).

    H <- H * n/(n+B); n <- n+B; Xs = sqrt(2/n) * fp32(x).reshape(-1, C); H <- H + Xs^T Xs
    """

    def __init__(self, columns):
        self.columns = int(columns)
        self.H = np.zeros((self.columns, self.columns), dtype=F32)
        self.nsamples = 0

    def add_batch(self, inp):
        inp = np.asarray(inp)
        if inp.ndim == 2:
            inp = inp[None]
        b = inp.shape[0]
        x = inp.reshape(-1, inp.shape[-1]).astype(F32)
        self.H = self.H * F32(self.nsamples / (self.nsamples + b))
        self.nsamples += b
        xs = (F32(math.sqrt(2 / self.nsamples)) * x).astype(F32)
        self.H = (self.H + xs.T @ xs).astype(F32)
        return self.H


def hessian_exact(batches):
    """(2/N) * sum x^T x in float64 (tolerance anchor for the tensor-core kernel)."""
    n = 0
    acc = None
    for inp in batches:
        inp = np.asarray(inp)
        if inp.ndim == 2:
            inp = inp[None]
        n += inp.shape[0]
        x = inp.reshape(-1, inp.shape[-1]).astype(np.float64)
        g = x.T @ x
        acc = g if acc is None else acc + g
    return acc * (2.0 / n)


# --------------------------------------------------------------------------------------
# A9/A10  SparseGPT OBS prune (fasterprune)
# --------------------------------------------------------------------------------------


def _chol_lower(H):
    """torch.linalg.cholesky with the reference's failure test (exception or NaN)."""
    try:
        with np.errstate(all="ignore"):
            L = np.linalg.cholesky(H.astype(F32))
    except np.linalg.LinAlgError:
        return None
    if np.isnan(L).any():
        return None
    return L.astype(F32)


def _fix_inf(H):
    """+inf -> 0.999 quantile, -inf -> 0.001 quantile -- sparsegpt_pruner.py:104-112,136-144."""
    pos = np.isposinf(H)
    if pos.any():
        H[pos] = np.quantile(H, 0.999)
    neg = np.isneginf(H)
    if neg.any():
        H[neg] = np.quantile(H, 0.001)
    return H


def obs_prepare_hinv(H, percdamp=0.01, max_retry=100):
    """fasterprune prologue -- sparsegpt_pruner.py:96-163.  Returns (Hinv_upper, dead_columns).

    NB the reference adds damping ONLY when a factorisation fails (unlike upstream SparseGPT)."""
    H = np.array(H, dtype=F32, copy=True)
    dead = np.diag(H) == 0
    H[dead, dead] = 1
    H = _fix_inf(H)
    damp = F32(percdamp) * np.mean(np.diag(H), dtype=F32)
    idx = np.arange(H.shape[0])
    for _ in range(max_retry):
        L = _chol_lower(H)
        if L is not None:
            break
        H[idx, idx] += damp
    else:
        raise RuntimeError("cholesky never succeeded")
    # torch.cholesky_inverse(L): (L L^T)^-1
    Linv = np.linalg.inv(L.astype(np.float64))
    Hi = (Linv.T @ Linv).astype(F32)
    Hi = _fix_inf(Hi)
    damp = F32(percdamp) * np.mean(np.abs(np.diag(Hi)), dtype=F32)
    for _ in range(max_retry):
        L2 = _chol_lower(Hi)
        if L2 is not None:
            break
        Hi[idx, idx] += damp
    else:
        raise RuntimeError("cholesky (upper) never succeeded")
    return np.ascontiguousarray(L2.T), dead


def obs_sweep(W, Hinv, sparsity, blocksize=128, prune_n=0, prune_m=0):
    """fasterprune block loop -- sparsegpt_pruner.py:172-213 (prune_n == 0: per-tile threshold; prune_n != 0: the n:m
    branch :195-198, decided column group by column group from the weights as updated so far, ties -> lower column).

    For each 128-column block: per-TILE threshold (same '<=' / +1 rule as the per-layer Wanda
    select), 128 sequential rank-1 updates inside the block, then the trailing update
    W[:, i2:] -= Err1 @ Hinv[i1:i2, i2:].  Returns (W_pruned, mask)."""
    W = np.array(W, dtype=F32, copy=True)
    R, C = W.shape
    mask_all = np.zeros((R, C), dtype=bool)
    for i1 in range(0, C, blocksize):
        i2 = min(i1 + blocksize, C)
        count = i2 - i1
        W1 = W[:, i1:i2].copy()
        Q1 = np.zeros_like(W1)
        Err1 = np.zeros_like(W1)
        Hinv1 = Hinv[i1:i2, i1:i2]
        d = np.diag(Hinv1).reshape(1, -1)
        if prune_n == 0:
            tmp = (W1 ** 2 / d ** 2).astype(F32)
            thresh = np.sort(tmp.reshape(-1), kind="stable")[int(tmp.size * sparsity)]
            mask1 = tmp <= thresh
        else:
            mask1 = np.zeros(W1.shape, dtype=bool)
        for i in range(count):
            if prune_n != 0 and i % prune_m == 0:
                tmp = (W1[:, i:i + prune_m] ** 2 / d[:, i:i + prune_m] ** 2).astype(F32)
                idx = np.argsort(tmp, axis=1, kind="stable")[:, :prune_n]  # topk(largest=False)
                np.put_along_axis(mask1[:, i:i + prune_m], idx, True, axis=1)
            w = W1[:, i]
            dd = Hinv1[i, i]
            q = w.copy()
            q[mask1[:, i]] = 0
            Q1[:, i] = q
            err1 = ((w - q) / dd).astype(F32)
            W1[:, i:] -= np.outer(err1, Hinv1[i, i:]).astype(F32)
            Err1[:, i] = err1
        W[:, i1:i2] = Q1
        mask_all[:, i1:i2] = mask1
        W[:, i2:] -= (Err1 @ Hinv[i1:i2, i2:]).astype(F32)
    return W, mask_all


def obs_prune(W, H, sparsity, blocksize=128, percdamp=0.01, out_dtype="fp32", prune_n=0, prune_m=0):
    """SparseGPT.fasterprune -- sparsegpt_pruner.py:84-218 (nn.Linear)."""
    W = np.array(W, dtype=F32, copy=True)
    Hinv, dead = obs_prepare_hinv(H, percdamp)
    W[:, dead] = 0
    Wp, mask = obs_sweep(W, Hinv, sparsity, blocksize, prune_n, prune_m)
    return round_to(Wp, out_dtype), mask


# --------------------------------------------------------------------------------------
# A11  zeroth-order perturbation
# --------------------------------------------------------------------------------------


def zo_perturb(W, z, scaling_factor, zo_eps, dtype):
    """param.data = param.data + scaling_factor * z * zo_eps
    -- layer_single_base_pruner.py:473-486.  Evaluated left to right, each op rounded to the
    parameter dtype: rn(w + rn(rn(scaling*z) * eps)).  z is drawn by the caller."""
    W = np.asarray(W, dtype=F32)
    z = np.asarray(z, dtype=F32)
    t = round_to(z * F32(scaling_factor), dtype)
    t = round_to(t * F32(zo_eps), dtype)
    return round_to(W + t, dtype)


def zo_projected_grad(loss1, loss2, zo_eps):
    """|(loss1 - loss2) / (2 * zo_eps)| -- layer_single_base_pruner.py:544-547 (python floats)."""
    return abs((loss1 - loss2) / (2 * zo_eps))


# --------------------------------------------------------------------------------------
# A12-A14  importance scores and group aggregation
# --------------------------------------------------------------------------------------


def mezo_importance(W, ghat, score_compute):
    """layer_single_base_pruner.py:551-559: GradOnly -> [ghat]; GradMagAbs -> |W|*ghat;
    GradMagSquare -> W^2 * ghat^2 (fp32)."""
    g = F32(ghat)
    if score_compute == "MEZO-GradOnly":
        return np.abs(np.array([g], dtype=F32))
    W = np.asarray(W, dtype=F32)
    if score_compute == "MEZO-GradMagAbs":
        return np.abs(W) * np.abs(g)
    if score_compute == "MEZO-GradMagSquare":
        return (W ** 2 * g ** 2).astype(F32)
    raise ValueError(score_compute)


def first_order_importance(W, gbar, score_compute):
    """layer_single_base_pruner.py:463-469 (gbar = mean over batches of |g| or g^2)."""
    W = np.asarray(W, dtype=F32)
    gbar = np.asarray(gbar, dtype=F32)
    if "GradMagSquare" in score_compute:
        return (W ** 2) * gbar
    if "GradMagAbs" in score_compute:
        return np.abs(W) * np.abs(gbar)
    if "GradOnly" in score_compute:
        return np.abs(gbar)
    raise ValueError(score_compute)


def abs_and_square_sums(W):
    """sum|w| and sum w^2 in float64 -- what the single segmented-reduction kernel produces per
    layer (A14: sum(|W| * ghat) == ghat * sum|W|,  sum(W^2 ghat^2) == ghat^2 * sum W^2)."""
    W = np.asarray(W, dtype=np.float64)
    return float(np.abs(W).sum()), float((W * W).sum())


def group_scores(importance_sums, numels, layer_to_group, aggregate):
    """return_sparsity group loop -- layer_single_base_pruner.py:342-377.
    importance_sums[l] = importance_measure[l].sum() (fp32).  Returns (scores, num_params)
    as insertion-ordered dicts."""
    scores, nparams = {}, {}
    for layer, group in layer_to_group.items():
        if group not in scores:
            scores[group] = F32(0)
            nparams[group] = 0
        scores[group] = F32(scores[group] + F32(importance_sums[layer]))
        nparams[group] += int(numels[layer])
    if aggregate == "avg":
        for g in scores:
            scores[g] = F32(scores[g] / F32(nparams[g]))
    return scores, nparams


# --------------------------------------------------------------------------------------
# A15  sparsity allocation
# --------------------------------------------------------------------------------------


def _sum_f32(x):
    """fp32 sum with float64 accumulation rounded once (torch's CPU cascade sum differs from this
    only by summation order; both are within 1 ulp of the exact sum for <= 1e3 terms)."""
    return F32(np.sum(np.asarray(x, dtype=np.float64)))


def sparsity_per_group(total_parameters_to_keep, group_scores_, group_num_parameters, max_sparsity_per_layer=0.8):
    """LayerSparsity.compute_the_sparsity_per_group -- layer_single_base_pruner.py:247-314.

    Reproduces the torch-CPU dtype promotions of the reference, op for op:
      * LongTensor * python float  -> fp32                              (:253, :299)
      * int64 + fp32 -> keep becomes fp32 after the first iteration      (:262)
      * the overshoot branch ADDS the removable amount (sic, :301)
    """
    keys = list(group_num_parameters.keys())
    scores = np.array([float(v) for v in group_scores_.values()], dtype=F32)
    nump = np.array([int(v) for v in group_num_parameters.values()], dtype=np.int64)
    one_minus = F32(1 - max_sparsity_per_layer)

    floor_keep = np.ceil(nump.astype(F32) * one_minus).astype(np.int32)
    keep = np.zeros(len(keys), dtype=np.int64) + floor_keep  # int64
    is_float = False

    def total(k):
        return _sum_f32(k) if is_float else int(k.sum())

    K = total_parameters_to_keep
    while total(keep) < K:
        total_ratio = _sum_f32(scores)
        if is_float:
            rest = F32(F32(K) - total(keep))  # python int - fp32 tensor -> fp32
        else:
            rest = K - int(keep.sum())  # python int - int64 tensor -> int64
        with np.errstate(all="ignore"):
            add = np.ceil((scores / total_ratio).astype(F32) * F32(rest)).astype(F32)
        keep = (keep.astype(F32) + add).astype(F32)
        is_float = True
        scores[keep >= nump.astype(F32)] = 0
        keep = np.minimum(keep, nump.astype(F32)).astype(F32)

        if _sum_f32(add) == 0:
            cur = total(keep)
            if cur < K:
                need = F32(F32(K) - cur)
                while need > 0:
                    for index in np.where(scores > 0)[0]:
                        can = min(need, F32(F32(nump[index]) - keep[index]))
                        keep[index] = F32(keep[index] + can)
                        need = F32(need - can)
                        if need == 0:
                            break
        if total(keep) > K:
            cur = total(keep)
            remove = F32(cur - F32(K))
            while remove > 0:
                order = np.argsort(-keep, kind="stable")
                for index in order:
                    floor_i = np.int32(F32(nump[index]) * one_minus)  # .int() truncation
                    can = min(remove, F32(keep[index] - F32(floor_i)))
                    keep[index] = F32(keep[index] + can)  # sic: '+=' in the reference (:301)
                    remove = F32(remove - can)
                    if remove == 0:
                        break

    out = {}
    for k, kk, n in zip(keys, keep, nump):
        if is_float:
            v = F32(1) - F32(kk) / F32(n)
        else:
            # int64 / int64 true division -> fp32 in torch
            v = F32(1) - F32(F32(kk) / F32(n))
        out[k] = float(min(max(F32(v), F32(0)), F32(1)))
    return out


def total_to_keep(total_parameters, original_sparsity):
    """int(total_parameters * (1 - original_sparsity)) -- layer_single_base_pruner.py:359."""
    return int(total_parameters * (1 - original_sparsity))


def layer_sparsity_from_scores(importance_sums, numels, layer_to_group, original_sparsity,
                               max_sparsity_per_layer, aggregate):
    """return_sparsity end to end (no prune_per_model) -- layer_single_base_pruner.py:342-414."""
    scores, nparams = group_scores(importance_sums, numels, layer_to_group, aggregate)
    K = total_to_keep(sum(int(numels[l]) for l in layer_to_group), original_sparsity)
    gs = sparsity_per_group(K, scores, nparams, max_sparsity_per_layer)
    return {layer: gs[g] for layer, g in layer_to_group.items()}


# --------------------------------------------------------------------------------------
# A16  grouping rules
# --------------------------------------------------------------------------------------


def block_group_name(name, n_fields):
    """'.'.join(name.split('.')[:n]) -- wanda_pruner.py:320 (T5: 4), :608 (ViT: 3), :766-768."""
    return ".".join(name.split(".")[:n_fields])


# --------------------------------------------------------------------------------------
# A17  sparsity check
# --------------------------------------------------------------------------------------


def count_zero(W):
    """(W == 0).sum() -- wanda_pruner.py:154."""
    return int((np.asarray(W) == 0).sum())


# --------------------------------------------------------------------------------------
# N3  global-pruner baselines -- oracle only so far: the CUDA path is a "next" row
# --------------------------------------------------------------------------------------
def global_iteration_ratios(target_sparsity, iterations):
    """p_i = target ** (iterations / i), i = 1..iterations -- global_pruner.py:162-164."""
    return [float(target_sparsity) ** (iterations / i) for i in range(1, iterations + 1)]


def global_get_mask(scores, p, max_sparsity_per_layer):
    """BLIPT5GlobalPruner.get_mask -- global_pruner.py:116-142.

    ``scores``: ordered dict name -> fp32 array.  Per layer the ``int(numel * (1 - max_sparsity))`` largest scores (and
    everything tied with the smallest of them, ``>=``) are protected by setting them to finfo.max; then ONE threshold --
    the ``int(p * total)``-th smallest of all (protected) scores -- keeps ``score > threshold`` everywhere.  Returns
    (masks as fp32 0/1 arrays, threshold).  ``int(p * total) == 0`` raises like ``threshold[-1]`` on an empty topk."""
    prot = {}
    for k, v in scores.items():
        v = np.array(v, dtype=F32, copy=True)
        num_to_set = int(v.size * (1 - max_sparsity_per_layer))
        if num_to_set > 0:
            t = np.sort(v.reshape(-1), kind="stable")[v.size - num_to_set]  # smallest of the num_to_set largest
            v[v >= t] = np.finfo(F32).max
        prot[k] = v
    allv = np.concatenate([v.reshape(-1) for v in prot.values()])
    num_to_zero_out = int(p * allv.size)
    if num_to_zero_out <= 0:
        raise IndexError("index -1 is out of bounds for dimension 0 with size 0")
    thres = np.sort(allv, kind="stable")[num_to_zero_out - 1]
    return {k: (v > thres).astype(F32) for k, v in prot.items()}, thres


def global_layerwise_mask(scores, p):
    """BLIPT5GlobalPruner.get_layerwise_mask -- global_pruner.py:144-157: the same rule per layer, no protection."""
    masks = {}
    for k, v in scores.items():
        v = np.asarray(v, dtype=F32)
        num_to_zero_out = int(p * v.size)
        if num_to_zero_out <= 0:
            raise IndexError("index -1 is out of bounds for dimension 0 with size 0")
        thres = np.sort(v.reshape(-1), kind="stable")[num_to_zero_out - 1]
        masks[k] = (v > thres).astype(F32)
    return masks
