"""Calibration accumulators with the reference's inner-seam interface, backed by the sm_100a kernels.

``WrappedGPT``  -- LAVIS/lavis/compression/pruners/wanda_pruner.py:54-84 (CoOp wanda_pruner.py:142-172,
                   UPop wanda_pruner.py:48-78): ``add_batch(inp, out)``, ``.scaler_row``, ``.nsamples``.
``SparseGPT``   -- LAVIS sparsegpt_pruner.py:56-222 (CoOp sparsegpt_pruner.py:145-311):
                   ``add_batch``, ``fasterprune(sparsity, prune_n, prune_m, blocksize, percdamp)``, ``free``.

Both call the C ABI through ``ecoflap_b200.ops``; there is no PyTorch implementation behind them.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import ops

try:  # only used for an isinstance check, exactly like the reference
    from transformers import Conv1D as _Conv1D
except Exception:  # pragma: no cover
    class _Conv1D:  # type: ignore
        pass


def _flatten_tokens(layer, inp):
    """[B, L, C] / [L, C] -> (batch entries B, [T, C] view).  2-D input counts as B = 1 (:72-74)."""
    if inp.dim() == 2:
        inp = inp.unsqueeze(0)
    b = inp.shape[0]
    if isinstance(layer, (nn.Linear, _Conv1D)) and inp.dim() == 3:
        inp = inp.reshape((-1, inp.shape[-1]))
    return b, inp


class NormBatch:
    """Collects ``WrappedGPT.add_batch`` calls and issues them as ONE kernel launch (``ecf_sqnorm_accum_batched``).
    The reference fires one hook per Linear per calibration batch (wanda_pruner.py:238-253); at BLIP-2 sizes a hook
    input is 1-25 MB, i.e. a few microseconds of HBM time, so per-hook launches are pure launch latency.  A batch may
    span the Linears of a block AND successive calibration batches (the kernel applies repeated updates of one
    accumulator in call order), so a block's whole calibration sweep becomes one launch of hundreds of MB.

    The hook inputs are only referenced, not copied, until ``flush()``; ``max_bytes`` bounds what is kept alive
    (an automatic flush happens when it is exceeded).  A block that overwrote a Linear's input in place between the
    hook and the flush would corrupt the norms, so the tensor version counters are checked and a mismatch raises
    (use ``WrappedGPT(...)`` without a batch for such a model)."""

    def __init__(self, max_bytes: int = 4 << 30):
        self._items = []   # (x, scaler_row, rescale, inv_n, version)
        self._bytes = 0
        self.max_bytes = int(max_bytes)

    def add(self, x, scaler_row, rescale, inv_n):
        self._items.append((x, scaler_row, rescale, inv_n, x._version))
        self._bytes += x.numel() * x.element_size()
        if self._bytes > self.max_bytes:
            self.flush()

    def flush(self):
        if not self._items:
            return
        for x, _, _, _, version in self._items:
            if x._version != version:
                raise RuntimeError("a hooked Linear input was modified in place before the batched norm launch; "
                                   "construct WrappedGPT without a NormBatch for this model")
        items, self._items, self._bytes = self._items, [], 0
        ops.sqnorm_accum_batched([it[:4] for it in items])

    def __len__(self):
        return len(self._items)


class HessianBatch:
    """Collects ``SparseGPT.add_batch`` calls of a block's calibration sweep and turns them into ONE tensor-core launch
    per distinct hook input (``ecf_hessian_accum`` with T = sum of the batches' tokens).

    Why: the LAVIS SparseGPT recipe hooks with batch size 1 (sparsegpt_pruner.py:390), i.e. T <= 257 tokens per call,
    and every call reads and writes all of H -- 8*C^2 bytes, 302 MB at C = 6144, ~46 us of HBM traffic for ~12 us of MMA.
    The running update  H <- H*n/(n+b) + (2/(n+b)) x^T x  telescopes to  H = (2/N) * X^T X  over the concatenated batches
    (same closed form NormBatch uses), so one launch over the concatenation gives the same matrix up to fp32 summation
    order, at the tensor-core rate instead of the H-traffic rate.  Linears that were fed the very same tensors in every
    batch (q/k/v, wi_0/wi_1, cross-attention k/v) have identical Hessians: the product is computed once, and those
    accumulators also share one inverse-Cholesky factor (``SparseGPT.prepare_hinv``) -- identical inputs, identical
    result, a third of the cuSOLVER work on a T5 block.

    The hook inputs are referenced, not copied, until ``flush()``; ``max_bytes`` bounds what is kept alive (an automatic
    flush folds what has been collected so far into the running H).  In-place modification of a referenced input is
    detected through the tensor version counter, as in NormBatch."""

    def __init__(self, max_bytes: int = 8 << 30):
        self._calls = {}    # id(acc) -> (acc, [(x, b, version)])
        self._bytes = 0
        self.max_bytes = int(max_bytes)

    def add(self, acc, x, b):
        self._calls.setdefault(id(acc), (acc, []))[1].append((x, b, x._version))
        self._bytes += x.numel() * x.element_size()
        if self._bytes > self.max_bytes:
            self.flush()

    def flush(self):
        if not self._calls:
            return
        calls, self._calls, self._bytes = self._calls, {}, 0
        groups = {}  # signature of the input tensors -> accumulators that saw exactly them
        for acc, items in calls.values():
            for x, _, version in items:
                if x._version != version:
                    raise RuntimeError("a hooked Linear input was modified in place before the batched Hessian launch; "
                                       "construct SparseGPT without a HessianBatch for this model")
            sig = (acc.nsamples, acc.columns) + tuple((x.data_ptr(), tuple(x.shape), tuple(x.stride()), x.dtype) for x, _, _ in items)
            groups.setdefault(sig, []).append((acc, items))
        for members in groups.values():
            acc0, items = members[0]
            xs = [x for x, _, _ in items]
            X = xs[0] if len(xs) == 1 else torch.cat(xs, dim=0)
            n_old = acc0.nsamples
            n_new = n_old + sum(b for _, b, _ in items)
            # H <- H * n_old/n_new + (2/n_new) X^T X : the closed form of the per-batch running update
            ops.hessian_accum(X, acc0.H, 2.0 / n_new, n_old / n_new)
            share = {} if len(members) > 1 else None
            for acc, _ in members:
                if acc is not acc0:
                    acc.H = acc0.H  # identical inputs: one matrix (read-only until prepare_hinv, which shares its result)
                acc._hinv_share = share
                acc.nsamples = n_new

    def __len__(self):
        return sum(len(v[1]) for v in self._calls.values())


# ECF_HINV_REFERENCE_ORDER=1 keeps the reference's three-step prologue for every matrix (A/B switch, parity debugging)
_REFERENCE_ORDER = os.environ.get("ECF_HINV_REFERENCE_ORDER", "0") == "1"
_EYES = {}  # (C, device) -> identity, the right-hand side of the triangular inverse


class WrappedGPT:
    """Running mean over samples of the per-input-channel sum of squared activations."""

    def __init__(self, layer, layer_id=0, layer_name="none", batch: "NormBatch | None" = None):
        self.layer = layer
        self.dev = self.layer.weight.device
        self.rows = layer.weight.data.shape[0]
        self.columns = layer.weight.data.shape[1]
        self.scaler_row = torch.zeros((self.columns), device=self.dev)
        self.nsamples = 0
        self.layer_id = layer_id
        self.layer_name = layer_name
        self.batch = batch  # optional NormBatch: the launch is deferred to batch.flush()

    def add_batch(self, inp, out=None):
        b, x = _flatten_tokens(self.layer, inp)
        n = self.nsamples + b
        # scaler_row = scaler_row * n_old/n + colsum(x^2) / n   -- one fused kernel, X read once
        if self.batch is not None:
            self.batch.add(x, self.scaler_row, self.nsamples / n, 1.0 / n)
        else:
            ops.sqnorm_accum(x, self.scaler_row, self.nsamples / n, 1.0 / n)
        self.nsamples = n


class SparseGPT:
    """Hessian accumulator + OBS pruning.

    ``H`` is kept as the running matrix the reference keeps (H*n/(n+b) + (2/n) X^T X) so that it can
    be inspected between batches; the rescale and the product are one tcgen05 kernel call.
    """

    def __init__(self, layer, batch: "HessianBatch | None" = None):
        self.layer = layer
        self.batch = batch          # optional HessianBatch: the launch is deferred to batch.flush()
        self._hinv_share = None     # set by HessianBatch for accumulators with identical inputs
        self.dev = self.layer.weight.device
        W = layer.weight.data
        if isinstance(self.layer, nn.Conv2d):
            W = W.flatten(1)
        if isinstance(self.layer, _Conv1D):
            W = W.t()
        self.rows = W.shape[0]
        self.columns = W.shape[1]
        self.H = torch.zeros((self.columns, self.columns), device=self.dev)
        self.nsamples = 0
        self.max_damp_retries = 64  # the reference retries forever (sparsegpt_pruner.py:117-131)

    def add_batch(self, inp, out=None):
        b, x = _flatten_tokens(self.layer, inp)
        if self.batch is not None:
            self.batch.add(self, x, b)  # nsamples is advanced by the flush
            return
        n = self.nsamples + b
        ops.hessian_accum(x, self.H, 2.0 / n, self.nsamples / n)
        self.nsamples = n

    # -- prologue helpers (cuSOLVER through torch.linalg) ---------------------------------
    @staticmethod
    def _repair_inf(H):
        pos = torch.isinf(H) & (H > 0)
        if pos.any():
            H[pos] = torch.quantile(H, 0.999)
        neg = torch.isinf(H) & (H < 0)
        if neg.any():
            H[neg] = torch.quantile(H, 0.001)
        return H

    def _cholesky_with_damping(self, H, damp, upper):
        """cholesky; on failure (error or NaN) add damp to the diagonal and retry -- damping is applied
        ONLY on failure, as in the reference (sparsegpt_pruner.py:117-131, 149-160)."""
        diag = torch.arange(self.columns, device=H.device)
        for _ in range(self.max_damp_retries):
            L, info = torch.linalg.cholesky_ex(H, upper=upper)
            if int(info.item()) == 0 and not torch.isnan(L).any():
                return L
            H[diag, diag] += damp
        raise RuntimeError("SparseGPT: Cholesky factorisation kept failing after damping retries")

    @staticmethod
    def _hinv_by_reversal(H):
        """The upper Cholesky factor U of H^-1 (H^-1 = U^T U) from ONE factorisation and ONE triangular inverse.

        With J the index reversal, J H J = L L^T gives H = V V^T with V = J L J upper triangular, hence
        H^-1 = V^-T V^-1 and U = V^-1 (upper, positive diagonal: the unique factor the reference reaches through
        cholesky -> cholesky_inverse -> cholesky(upper=True), sparsegpt_pruner.py:117-160).  Measured on B200 for
        C = 6144: 11.2 ms against 27.1 ms, and closer to the fp64 factor than the three-step order
        (profiles/r3/chol_probe.log).  One host sync.  Returns None -- the caller then runs the reference's own order
        with its inf repair and failure-only damping -- when H holds a non-finite entry, is not positive definite, or
        the inverse is not finite."""
        C = H.shape[0]
        Lf, info = torch.linalg.cholesky_ex(H.flip(0, 1))
        key = (C, H.device)
        eye = _EYES.get(key)
        if eye is None:
            eye = _EYES[key] = torch.eye(C, device=H.device, dtype=H.dtype)
        # (V^T)^-1 = U^T; the solver answers in column-major order, so the transpose is U, row-major, without a copy
        U = torch.linalg.solve_triangular(Lf.flip(0, 1).T, eye, upper=False).T
        if not U.is_contiguous():
            U = U.contiguous()
        ok = torch.isfinite(H).all() & (info == 0) & torch.isfinite(U).all()
        return U if bool(ok.item()) else None

    def prepare_hinv(self, percdamp=0.01):
        """Returns (Hinv_upper, dead_mask) and releases H (sparsegpt_pruner.py:96-163)."""
        share = self._hinv_share
        if share is not None and "hinv" in share:  # a Linear with the very same inputs has factorised this H already
            self.H = None
            return share["hinv"]
        H = self.H
        del self.H
        dead = torch.diag(H) == 0
        H[dead, dead] = 1
        if not _REFERENCE_ORDER:
            U = self._hinv_by_reversal(H)
            if U is not None:
                res = (U, dead)
                if share is not None:
                    share["hinv"] = res
                return res
        H = self._repair_inf(H)
        damp = percdamp * torch.mean(torch.diag(H))
        L = self._cholesky_with_damping(H, damp, upper=False)
        Hi = torch.cholesky_inverse(L)
        Hi = self._repair_inf(Hi)
        damp = percdamp * torch.mean(torch.diag(Hi).abs())
        Hinv = self._cholesky_with_damping(Hi, damp, upper=True)
        res = (Hinv.contiguous(), dead)
        if share is not None:
            share["hinv"] = res
        return res

    def fasterprune(self, sparsity, prune_n=0, prune_m=0, blocksize=128, percdamp=0.01):
        W = self.layer.weight.data.clone()
        if isinstance(self.layer, nn.Conv2d):
            W = W.flatten(1)
        if isinstance(self.layer, _Conv1D):
            W = W.t()
        W = W.float().contiguous()
        Hinv, dead = self.prepare_hinv(percdamp)
        W[:, dead] = 0
        kth = []
        for i1 in range(0, self.columns, blocksize):
            count = min(i1 + blocksize, self.columns) - i1
            kth.append(int(self.rows * count * sparsity))  # int(tmp.numel() * sparsity), :187
        ops.obs_prune(W, Hinv, kth, blocksize, prune_n, prune_m)
        if isinstance(self.layer, _Conv1D):
            W = W.t()
        # in place (the reference rebinds .data, sparsegpt_pruner.py:214): same values, but the storage the block's captured
        # forward graph reads (pruners/sweep.py, _BlockReplay) stays the one that holds the pruned weights
        self.layer.weight.data.copy_(W.reshape(self.layer.weight.shape))

    def free(self):
        # the reference also calls torch.cuda.empty_cache() here (sparsegpt_pruner.py:220-222): 588 cudaFree / cudaMalloc
        # round trips per BLIP-2 run that only make the next H allocation slower -- the block stays in torch's caching
        # allocator instead
        self.H = None


__all__ = ["NormBatch", "HessianBatch", "WrappedGPT", "SparseGPT"]
