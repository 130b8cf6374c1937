"""Torch-tensor facing wrappers over the C ABI: pointer / stream / workspace plumbing only.

Every function here launches the hand-written sm_100a kernels through ``_abi.lib``; nothing in this
module computes with PyTorch ops, and there is no CPU path (CPU tensors raise).
"""
from __future__ import annotations

import ctypes

import torch

from . import _abi
from ._abi import check, lib

_DT = {torch.float32: _abi.ECF_F32, torch.float16: _abi.ECF_F16, torch.bfloat16: _abi.ECF_BF16}


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"ecoflap_b200 supports fp32/fp16/bf16 tensors, got {t.dtype}") from None


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("ecoflap_b200 kernels need CUDA tensors (there is no CPU fallback)")


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


class _Workspaces:
    """One growing scratch buffer per (device, stream): calls on one stream are serialised, so a single
    buffer can be shared by all ops issued on it."""

    def __init__(self):
        self._bufs = {}

    def get(self, device, nbytes: int, tag: str = "") -> torch.Tensor:
        # one buffer per op family: some kernels keep self-resetting counters at the start of their workspace
        key = (device.index, torch.cuda.current_stream(device).cuda_stream, tag)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.zeros(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
            self._bufs[key] = buf
        return buf


_ws = _Workspaces()


def _as_2d(x: torch.Tensor):
    """[..., C] -> (base tensor, T, C, ld) without copying when the rows are unit-stride."""
    C = x.shape[-1]
    x2 = x.reshape(-1, C)
    if x2.stride(1) != 1 or (x2.shape[0] > 1 and x2.stride(0) < C):
        x2 = x2.contiguous()
    ld = x2.stride(0) if x2.shape[0] > 1 else C
    return x2, x2.shape[0], C, ld


def sqnorm_accum(x: torch.Tensor, scaler_row: torch.Tensor, rescale: float, inv_n: float) -> None:
    """scaler_row = scaler_row * rescale + colsum(x^2) * inv_n   (A1, wanda_pruner.py:71-84)."""
    _require_cuda(x, scaler_row)
    assert scaler_row.dtype == torch.float32 and scaler_row.is_contiguous()
    x2, T, C, ld = _as_2d(x)
    assert scaler_row.numel() == C
    need = lib.ecf_workspace_bytes(_abi.OP_SQNORM, T, C)
    ws = _ws.get(x.device, need, "sqnorm")
    check(lib.ecf_sqnorm_accum(x2.data_ptr(), dtype_code(x2), T, C, ld, scaler_row.data_ptr(), float(rescale),
                               float(inv_n), ws.data_ptr(), ws.numel(), _stream(x)))


def sqnorm_accum_batched(items) -> None:
    """One launch for a list of ``(x, scaler_row, rescale, inv_n)`` hook calls (A1, batched).  ``x`` tensors may be
    shared between items (q/k/v see the same input); items may also repeat a ``scaler_row`` (successive calibration
    batches of one Linear), in which case they are applied in list order.  Lists beyond the ABI's limits (256 calls /
    32 accumulators per launch) are issued as several launches, still in order."""
    items = list(items)
    keep = []  # contiguous copies must outlive the launch call
    i = 0
    while i < len(items):
        rows, j = set(), i
        while j < len(items) and j - i < _abi.SQNORM_MAX_BATCH:
            ptr = items[j][1].data_ptr()
            if ptr not in rows and len(rows) == _abi.SQNORM_MAX_GROUPS:
                break
            rows.add(ptr)
            j += 1
        chunk = items[i:j]
        i = j
        descs = (_abi.SqnormDesc * len(chunk))()
        for d, (x, scaler_row, rescale, inv_n) in zip(descs, chunk):
            _require_cuda(x, scaler_row)
            assert scaler_row.dtype == torch.float32 and scaler_row.is_contiguous()
            x2, T, C, ld = _as_2d(x)
            assert scaler_row.numel() == C
            keep.append(x2)
            d.x, d.scaler_row, d.T, d.C, d.ld = x2.data_ptr(), scaler_row.data_ptr(), T, C, ld
            d.dtype, d.rescale, d.inv_n = dtype_code(x2), float(rescale), float(inv_n)
        dev = chunk[0][0].device
        need = lib.ecf_sqnorm_batched_workspace_bytes(descs, len(chunk))
        ws = _ws.get(dev, need, "sqnorm")
        check(lib.ecf_sqnorm_accum_batched(descs, len(chunk), ws.data_ptr(), ws.numel(), _stream(chunk[0][0])))


def _weight_2d(W: torch.Tensor):
    assert W.dim() == 2 and W.stride(1) == 1, "weight must be a row-major 2-D tensor"
    return W.shape[0], W.shape[1], (W.stride(0) if W.shape[0] > 1 else W.shape[1])


def alloc_mask_bits(R: int, C: int, device) -> torch.Tensor:
    return torch.zeros((R, (C + 7) // 8), dtype=torch.uint8, device=device)


def unpack_mask_bits(mask_bits: torch.Tensor, C: int) -> torch.Tensor:
    """Packed mask -> bool [R, C] (test/debug helper; bit j of byte v is column 8v+j)."""
    shifts = torch.arange(8, device=mask_bits.device, dtype=torch.uint8)
    bits = (mask_bits.unsqueeze(-1) >> shifts) & 1
    return bits.reshape(mask_bits.shape[0], -1)[:, :C].bool()


def wanda_row_select_apply(W, scaler_row, k_per_row: int, mask_bits=None, n_zero=None) -> None:
    """Zero, in place, the k smallest |W|*sqrt(scaler_row) of every row (A3+A4+A7)."""
    _require_cuda(W, scaler_row, mask_bits, n_zero)
    R, C, ld = _weight_2d(W)
    assert scaler_row.dtype == torch.float32 and scaler_row.numel() == C and scaler_row.is_contiguous()
    check(lib.ecf_wanda_row_select_apply(
        W.data_ptr(), dtype_code(W), R, C, ld, scaler_row.data_ptr(), int(k_per_row),
        mask_bits.data_ptr() if mask_bits is not None else None,
        mask_bits.stride(0) if mask_bits is not None else 0,
        n_zero.data_ptr() if n_zero is not None else None, None, 0, _stream(W)))


def wanda_nm_select_apply(W, scaler_row, n: int, m: int, mask_bits=None, n_zero=None) -> None:
    """n:m structured select (A6): zero, in place, the n smallest |W|*sqrt(scaler_row) of every group of m consecutive
    columns.  Raises like ``torch.topk`` when a group is shorter than n."""
    _require_cuda(W, scaler_row, mask_bits, n_zero)
    R, C, ld = _weight_2d(W)
    assert scaler_row.dtype == torch.float32 and scaler_row.numel() == C and scaler_row.is_contiguous()
    last = C % m if C % m else m
    if n > m or n > last:
        raise RuntimeError("selected index k out of range")  # torch.topk's message
    check(lib.ecf_wanda_nm_select_apply(
        W.data_ptr(), dtype_code(W), R, C, ld, scaler_row.data_ptr(), int(n), int(m),
        mask_bits.data_ptr() if mask_bits is not None else None,
        mask_bits.stride(0) if mask_bits is not None else 0,
        n_zero.data_ptr() if n_zero is not None else None, _stream(W)))


def wanda_row_select_apply_batched(items) -> None:
    """Per-row select of several matrices (the Linears of one block): matrices with the same row length and dtype
    share one persistent launch.  ``items`` is a list of ``(W, scaler_row, k_per_row)`` or
    ``(W, scaler_row, k_per_row, mask_bits, n_zero)`` tuples; lists longer than the ABI's batch limit are issued as
    several calls."""
    items = [tuple(it) + (None,) * (5 - len(it)) for it in items]
    for start in range(0, len(items), _abi.ROW_MAX_BATCH):
        chunk = items[start:start + _abi.ROW_MAX_BATCH]
        descs = (_abi.RowDesc * len(chunk))()
        for d, (W, scaler_row, k, mask_bits, n_zero) in zip(descs, chunk):
            _require_cuda(W, scaler_row, mask_bits, n_zero)
            R, C, ld = _weight_2d(W)
            assert scaler_row.dtype == torch.float32 and scaler_row.numel() == C and scaler_row.is_contiguous()
            d.W, d.scaler_row, d.R, d.C, d.ld, d.dtype, d.k_per_row = W.data_ptr(), scaler_row.data_ptr(), R, C, ld, dtype_code(W), int(k)
            d.mask_bits = mask_bits.data_ptr() if mask_bits is not None else None
            d.mask_ld = mask_bits.stride(0) if mask_bits is not None else 0
            d.n_zero = n_zero.data_ptr() if n_zero is not None else None
        check(lib.ecf_wanda_row_select_apply_batched(descs, len(chunk), None, 0, _stream(chunk[0][0])))


def wanda_layer_thresh_apply(W, scaler_row, kth_index: int, thres_out=None, mask_bits=None, n_zero=None) -> None:
    """Zero, in place, every entry whose score is <= the kth_index-th smallest score (A3+A5+A7)."""
    _require_cuda(W, scaler_row, mask_bits, n_zero, thres_out)
    R, C, ld = _weight_2d(W)
    assert scaler_row.dtype == torch.float32 and scaler_row.numel() == C and scaler_row.is_contiguous()
    if not (0 <= kth_index < R * C):
        raise IndexError(f"index {kth_index} is out of bounds for dimension 0 with size {R * C}")
    need = lib.ecf_workspace_bytes(_abi.OP_LAYER_THRESH, R, C)
    ws = _ws.get(W.device, need, "layer_thresh")
    check(lib.ecf_wanda_layer_thresh_apply(
        W.data_ptr(), dtype_code(W), R, C, ld, scaler_row.data_ptr(), int(kth_index),
        thres_out.data_ptr() if thres_out is not None else None,
        mask_bits.data_ptr() if mask_bits is not None else None,
        mask_bits.stride(0) if mask_bits is not None else 0,
        n_zero.data_ptr() if n_zero is not None else None, ws.data_ptr(), ws.numel(), _stream(W)))


def wanda_layer_thresh_apply_batched(items) -> None:
    """Per-layer select of several matrices (the Linears of one block) in one cooperative launch.  ``items`` is a list
    of ``(W, scaler_row, kth_index)`` or ``(W, scaler_row, kth_index, thres_out, mask_bits, n_zero)`` tuples; lists
    longer than the ABI's batch limit are issued as several launches."""
    items = [tuple(it) + (None,) * (6 - len(it)) for it in items]
    for start in range(0, len(items), _abi.LAYER_MAX_BATCH):
        chunk = items[start:start + _abi.LAYER_MAX_BATCH]
        descs = (_abi.LayerDesc * len(chunk))()
        for d, (W, scaler_row, kth, thres_out, mask_bits, n_zero) in zip(descs, chunk):
            _require_cuda(W, scaler_row, mask_bits, n_zero, thres_out)
            R, C, ld = _weight_2d(W)
            assert scaler_row.dtype == torch.float32 and scaler_row.numel() == C and scaler_row.is_contiguous()
            if not (0 <= kth < R * C):
                raise IndexError(f"index {kth} is out of bounds for dimension 0 with size {R * C}")
            d.W, d.scaler_row, d.R, d.C, d.ld, d.dtype, d.kth_index = W.data_ptr(), scaler_row.data_ptr(), R, C, ld, dtype_code(W), int(kth)
            d.thres_out = thres_out.data_ptr() if thres_out is not None else None
            d.mask_bits = mask_bits.data_ptr() if mask_bits is not None else None
            d.mask_ld = mask_bits.stride(0) if mask_bits is not None else 0
            d.n_zero = n_zero.data_ptr() if n_zero is not None else None
        dev = chunk[0][0].device
        need = lib.ecf_layer_thresh_batched_workspace_bytes(descs, len(chunk))
        ws = _ws.get(dev, need, "layer_thresh")
        check(lib.ecf_wanda_layer_thresh_apply_batched(descs, len(chunk), ws.data_ptr(), ws.numel(), _stream(chunk[0][0])))


def layer_thresh_phase_times_us(device) -> list:
    """Profiling aid: durations (us) of the phases P1..P5 of the LAST batched per-layer select on the current stream
    of ``device`` (read from the %globaltimer stamps the kernel leaves at the start of its workspace).  Synchronises."""
    torch.cuda.synchronize(device)
    ws = _ws.get(device, 256, "layer_thresh")
    t = ws[:48].view(torch.int64).tolist()
    return [(b - a) / 1e3 for a, b in zip(t[:-1], t[1:])]


def layer_thresh_last_fallback(device) -> bool:
    """Test / profiling aid: did the LAST batched per-layer select on ``device`` leave the split path and run the
    cooperative kernel (k-th score outside the sampled bracket, heavy ties, a full bracket list)?  Synchronises."""
    torch.cuda.synchronize(device)
    ws = _ws.get(device, 8192, "layer_thresh")
    off = int(lib.ecf_layer_thresh_flag_offset())  # the "flag of the last launch" word of the workspace header
    return bool(ws[off:off + 4].view(torch.int32).item() != 0)


def layer_thresh_stamps_us(device) -> list:
    """Profiling aid: all 16 %globaltimer stamps of CTA 0 (us, relative to the kernel start; 0 = not taken)."""
    torch.cuda.synchronize(device)
    ws = _ws.get(device, 256, "layer_thresh")
    t = ws[:128].view(torch.int64).tolist()
    return [round((x - t[0]) / 1e3, 1) if x else 0 for x in t]


def zo_perturb(W: torch.Tensor, z: torch.Tensor, scaling: float, eps: float) -> None:
    """W = rn(W + rn(rn(scaling*z)*eps)) in W's dtype, in place (A11)."""
    _require_cuda(W, z)
    assert W.is_contiguous() and z.is_contiguous() and W.dtype == z.dtype and W.numel() == z.numel()
    check(lib.ecf_zo_perturb(W.data_ptr(), dtype_code(W), W.numel(), z.data_ptr(), float(scaling), float(eps),
                             _stream(W)))


def count_zero(W: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """out(uint64 as int64 tensor) += number of zeros in W (A17)."""
    _require_cuda(W)
    Wc = W if W.is_contiguous() else W.contiguous()
    if out is None:
        out = torch.zeros(1, dtype=torch.int64, device=W.device)
    check(lib.ecf_count_zero(Wc.data_ptr(), dtype_code(Wc), Wc.numel(), out.data_ptr(), _stream(W)))
    return out


def group_abs_reduce(tensors):
    """Per tensor: (sum|w|, sum w^2) as two float64 CUDA tensors, one launch for the whole list (A14)."""
    tensors = [t if t.is_contiguous() else t.contiguous() for t in tensors]
    _require_cuda(*tensors)
    dev = tensors[0].device
    chunk = lib.ecf_group_reduce_chunk_elems()
    n = len(tensors)
    table = (_abi.TensorDesc * n)()
    cb = 0
    for i, t in enumerate(tensors):
        table[i].ptr = t.data_ptr()
        table[i].numel = t.numel()
        table[i].dtype = dtype_code(t)
        table[i].reserved = 0
        table[i].chunk_begin = cb
        cb += max(1, (t.numel() + chunk - 1) // chunk)
    raw = bytes(table)
    host = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
    d_table = host.to(dev, non_blocking=False)
    sum_abs = torch.zeros(n, dtype=torch.float64, device=dev)
    sum_sq = torch.zeros(n, dtype=torch.float64, device=dev)
    need = lib.ecf_workspace_bytes(_abi.OP_GROUP_REDUCE, n, cb)
    ws = _ws.get(dev, need, "group_reduce")
    check(lib.ecf_group_abs_reduce(d_table.data_ptr(), n, cb, sum_abs.data_ptr(), sum_sq.data_ptr(), ws.data_ptr(),
                                   ws.numel(), _stream(tensors[0])))
    return sum_abs, sum_sq


GLOBAL_MODES = {"mag": _abi.GLOBAL_MAG, "grad_mag_abs": _abi.GLOBAL_GRAD_MAG_ABS, "grad_mag_sq": _abi.GLOBAL_GRAD_MAG_SQ,
                "grad_only": _abi.GLOBAL_GRAD_ONLY}


class GlobalTable:
    """Device table of the tensors a global-pruner step works on (N3; global_pruner.py:116-207).  ``weights``: contiguous
    parameter tensors pruned in place; ``grads``: matching fp32 sums of |grad| (grad^2) over ``n_batches`` batches, or None
    for the magnitude score."""

    def __init__(self, weights, grads=None, n_batches=1, mode="mag"):
        _require_cuda(*weights)
        assert all(w.is_contiguous() for w in weights), "global select: parameters must be contiguous"
        self.weights, self.grads = list(weights), grads
        self.mode, self.n_batches = GLOBAL_MODES[mode], float(n_batches)
        self.dev = weights[0].device
        chunk = lib.ecf_global_chunk_elems()
        n = len(weights)
        table = (_abi.GlobalDesc * n)()
        cb = 0
        for i, w in enumerate(weights):
            g = None if grads is None else grads[i]
            if g is not None:
                assert g.dtype == torch.float32 and g.is_contiguous() and g.numel() == w.numel() and g.device == w.device
            else:
                assert self.mode == _abi.GLOBAL_MAG, "this score mode needs the accumulated gradients"
            assert w.numel() < (1 << 32), "global select: tensors beyond 2^32 elements are not supported"
            table[i].W = w.data_ptr()
            table[i].G = 0 if g is None else g.data_ptr()
            table[i].numel = w.numel()
            table[i].dtype = dtype_code(w)
            table[i].reserved = 0
            table[i].chunk_begin = cb
            cb += max(1, (w.numel() + chunk - 1) // chunk)
        self.n, self.chunks = n, cb
        self.numels = [w.numel() for w in weights]
        self.d_table = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).to(self.dev)

    def select(self, ranks, segmented, protect=None):
        """0-based ascending ranks (one per segment) -> uint32 threshold keys on the device (int64 tensor view of uint32)."""
        nseg = self.n if segmented else 1
        assert len(ranks) == nseg
        d_ranks = torch.tensor([int(r) for r in ranks], dtype=torch.int64, device=self.dev)
        tkeys = torch.zeros(nseg, dtype=torch.int32, device=self.dev)
        need = lib.ecf_workspace_bytes(_abi.OP_GLOBAL_SELECT, nseg, 0)
        ws = _ws.get(self.dev, need, "global_select")
        check(lib.ecf_global_select(self.d_table.data_ptr(), self.n, self.chunks, self.mode, self.n_batches, int(bool(segmented)),
                                    0 if protect is None else protect.data_ptr(), d_ranks.data_ptr(), tkeys.data_ptr(),
                                    ws.data_ptr(), ws.numel(), _stream(self.weights[0])))
        return tkeys

    def score_sums(self):
        """Per-tensor sum of the per-element score as a float64 CUDA tensor (one launch for the whole table)."""
        sums = torch.zeros(self.n, dtype=torch.float64, device=self.dev)
        check(lib.ecf_global_score_sum(self.d_table.data_ptr(), self.n, self.chunks, self.mode, self.n_batches, sums.data_ptr(),
                                       _stream(self.weights[0])))
        return sums

    def apply(self, tkeys, segmented, protect=None):
        """w *= (score > threshold) in place; returns the per-tensor number of elements at or below the threshold."""
        pruned = torch.zeros(self.n, dtype=torch.int64, device=self.dev)
        check(lib.ecf_global_apply(self.d_table.data_ptr(), self.n, self.chunks, self.mode, self.n_batches, int(bool(segmented)),
                                   0 if protect is None else protect.data_ptr(), tkeys.data_ptr(), pruned.data_ptr(),
                                   _stream(self.weights[0])))
        return pruned


def grad_accum(G: torch.Tensor, g: torch.Tensor, square: bool = False) -> None:
    """G += |g| (or g^2) in place, G fp32 (A13: the first-order loops' gradient accumulator, kept on the device)."""
    _require_cuda(G, g)
    assert G.dtype == torch.float32 and G.is_contiguous() and G.numel() == g.numel()
    g = g if g.is_contiguous() else g.contiguous()
    check(lib.ecf_grad_accum(G.data_ptr(), g.data_ptr(), dtype_code(g), g.numel(), int(bool(square)), _stream(G)))


def global_key_to_float(key: int) -> float:
    """Inverse of the kernels' order-preserving key (for reporting the threshold)."""
    import struct

    key &= 0xFFFFFFFF
    bits = (key & 0x7FFFFFFF) if key & 0x80000000 else (~key & 0xFFFFFFFF)
    return struct.unpack("<f", struct.pack("<I", bits))[0]


def hessian_accum(x: torch.Tensor, H: torch.Tensor, alpha: float, beta: float) -> None:
    """H = beta*H + alpha * x^T x  on tcgen05 tensor cores (A8, sparsegpt_pruner.py:71-82)."""
    _require_cuda(x, H)
    assert H.dtype == torch.float32 and H.dim() == 2 and H.stride(1) == 1
    x2, T, C, ld = _as_2d(x)
    assert H.shape[0] == C and H.shape[1] == C
    need = lib.ecf_workspace_bytes(_abi.OP_HESSIAN, T if x2.dtype == torch.float32 else 0, C)
    ws = _ws.get(x.device, need, "hessian")
    check(lib.ecf_hessian_accum(x2.data_ptr(), dtype_code(x2), T, C, ld, H.data_ptr(), H.stride(0), float(alpha),
                                float(beta), ws.data_ptr(), ws.numel(), _stream(x)))


def obs_prune(W: torch.Tensor, Hinv: torch.Tensor, kth_per_block, blocksize: int = 128, prune_n: int = 0, prune_m: int = 0) -> None:
    """SparseGPT block loop on an fp32 working copy W, in place (A10, sparsegpt_pruner.py:172-213); prune_n != 0 selects the
    n:m branch (:182-198), where ``kth_per_block`` is ignored."""
    _require_cuda(W, Hinv)
    assert W.dtype == torch.float32 and Hinv.dtype == torch.float32
    R, C, ldw = _weight_2d(W)
    assert Hinv.shape == (C, C) and Hinv.stride(1) == 1
    nb = (C + blocksize - 1) // blocksize
    if prune_n != 0:
        kth_per_block = [0] * nb
    assert len(kth_per_block) == nb
    arr = (ctypes.c_int64 * nb)(*[int(v) for v in kth_per_block])
    need = lib.ecf_workspace_bytes(_abi.OP_OBS, R, C)
    ws = _ws.get(W.device, need, "obs")
    check(lib.ecf_obs_prune(W.data_ptr(), R, C, ldw, Hinv.data_ptr(), Hinv.stride(0), arr, int(blocksize), int(prune_n),
                            int(prune_m), ws.data_ptr(), ws.numel(), _stream(W)))
