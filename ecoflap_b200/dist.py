"""Multi-GPU partitioning of the hot path (new design -- the reference runs on one GPU).

One process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch on the box, gloo in the CPU tests):

  * calibration batches are sharded round-robin over the ranks; every rank accumulates the running mean of
    sum x^2 (or the Hessian) over ITS samples; one all-reduce per block over the concatenated per-Linear
    vectors turns them into the global mean  sum_r n_r * s_r / sum_r n_r  (latency bound: <= 35 k floats);
  * per-ROW mask selection can be sharded over output rows (rows are independent): rank r selects rows
    [r*R/P, (r+1)*R/P) in place and an all-gather rebuilds the pruned matrix on every rank.  Measured on B200 the
    all-gather (NVLink, ~0.9 TB/s per direction) costs more than it saves against the replicated HBM-bound select
    (> 2 TB/s of algorithmic bytes), so replication is the default and sharding an option (bench: ECF_ROW_SHARD=1);
  * the per-LAYER threshold select and the OBS sweep are replicated (their exchange steps are future work).

The collective plumbing below works on CPU and CUDA tensors alike; the kernels are only reached through the
``select_rows`` callable, so the world-size-2 gloo tests exercise exactly this code.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch
import torch.distributed as dist


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank_world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def zo_draws_per_layer(batch_lens: Sequence[int], num_samples: int, num_noise: int) -> int:
    """How many (batch, noise) evaluations -- i.e. ``np.random.randint`` seed draws -- the reference's zeroth-order loop
    makes for ONE parameter (layer_single_base_pruner.py:515-547): ``seen`` grows by the batch length per noise draw
    and both loops stop once it reaches ``num_samples``.  ``batch_lens``: batch lengths in loader order."""
    seen, draws = 0, 0
    for bl in batch_lens:
        if seen >= num_samples:
            break
        for _ in range(num_noise):
            if seen >= num_samples:
                break
            seen += int(bl)
            draws += 1
    return draws


def zo_owner(layer_index: int, world: int) -> int:
    """Layer -> rank assignment of the sharded zeroth-order loop: l = r (mod P)."""
    return layer_index % world


def shard_indices(n: int, rank: int, world: int):
    """Round-robin ownership of calibration batches: j = rank (mod world)."""
    return list(range(rank, n, world))


def row_range(rows: int, rank: int, world: int):
    """Contiguous, equal row shards (requires rows % world == 0 for the in-place all-gather)."""
    per = rows // world
    return rank * per, (rank + 1) * per


def allreduce_running_means(stats: Sequence[torch.Tensor], counts: Sequence[int], group=None, totals: Sequence[int] = None):
    """In place: every tensor in ``stats`` (a per-rank running mean over ``counts[i]`` samples) becomes the
    mean over the samples of all ranks.  One all-reduce for the whole list.  Returns the global counts.

    ``totals``: the global sample counts when the caller already knows them (equal shards: count * world).  The
    counts then stay on the host, nothing is read back from the device and the call can be captured in a CUDA graph;
    without it the counts ride along in the all-reduce and are read back (one host sync)."""
    if not stats:
        return []
    flat = torch.cat([(s.reshape(-1) * float(n)) for s, n in zip(stats, counts)])
    if totals is not None:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for s, tot in zip(stats, totals):
            n = s.numel()
            s.copy_((flat[off:off + n] * (1.0 / float(tot))).reshape(s.shape))
            off += n
        return [int(t) for t in totals]
    tail = torch.tensor([float(n) for n in counts], dtype=flat.dtype, device=flat.device)
    buf = torch.cat([flat, tail])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    totals_dev = buf[flat.numel():]
    off = 0
    out_counts = []
    for i, s in enumerate(stats):
        n = s.numel()
        tot = totals_dev[i]
        s.copy_((buf[off:off + n] / tot).reshape(s.shape))
        off += n
        out_counts.append(int(round(float(tot))))
    return out_counts


def sync_block_norms(accumulators, group=None, totals: Sequence[int] = None):
    """accumulators: WrappedGPT-like objects (.scaler_row fp32 [C], .nsamples)."""
    accs = list(accumulators)
    totals = allreduce_running_means([a.scaler_row for a in accs], [a.nsamples for a in accs], group, totals)
    for a, n in zip(accs, totals):
        a.nsamples = n


def pack_block_norms(accumulators) -> torch.Tensor:
    """Rebind the (still empty) norm accumulators of a block to slices of ONE flat fp32 buffer, so that the block's
    exchange step is a single in-place all-reduce with no gather / scatter kernels around it.  Returns the buffer."""
    accs = list(accumulators)
    flat = torch.zeros(sum(a.columns for a in accs), dtype=torch.float32, device=accs[0].dev)
    off = 0
    for a in accs:
        assert a.nsamples == 0, "pack_block_norms must run before the first add_batch"
        a.scaler_row = flat[off:off + a.columns]
        off += a.columns
    return flat


def sync_packed_norms(flat: torch.Tensor, accumulators, group=None):
    """Exchange step for accumulators packed by ``pack_block_norms`` when every rank holds the same number of samples:
    the global running mean is the average of the per-rank means -- one all-reduce (NCCL: ReduceOp.AVG), in place."""
    _, world = rank_world(group)
    if world > 1:
        if dist.get_backend(group) == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            flat.mul_(1.0 / world)
    for a in accumulators:
        a.nsamples *= world


class PeerNormExchange:
    """Exchange step of the norm accumulators over NVSwitch peer memory (``ecf_norm_exchange_p2p``): one small fused
    copy / signal / gather-sum kernel per rank and block instead of an NCCL all-reduce.  torch's symmetric-memory
    rendezvous supplies the peer mappings of the staging buffer (plumbing); the exchange itself is our kernel.

    ``available()`` is False off-GPU, for non-NCCL groups, or when the rendezvous is not supported; callers then use
    ``sync_packed_norms`` (NCCL)."""

    def __init__(self, max_floats: int, device, group=None):
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem

        from . import _abi

        self._abi = _abi
        self.rank, self.world = rank_world(group)
        self.max_floats = (int(max_floats) + 3) // 4 * 4
        nbytes = _abi.lib.ecf_norm_exchange_staging_bytes(self.max_floats)
        self.staging = symm_mem.empty((nbytes + 3) // 4, dtype=torch.float32, device=device)
        self.staging.zero_()
        torch.cuda.synchronize(device)
        grp = group if group is not None else dist.group.WORLD
        self.handle = symm_mem.rendezvous(self.staging, grp)
        dist.barrier(group)  # every rank's flags are zero before anybody signals
        ptrs = list(self.handle.buffer_ptrs)
        assert len(ptrs) == self.world and self.handle.rank == self.rank
        self._ptrs = (ctypes.c_void_p * self.world)(*ptrs)

    @staticmethod
    def available(group=None) -> bool:
        try:
            if not (torch.cuda.is_available() and is_dist() and str(dist.get_backend(group)) == "nccl"):
                return False
            import torch.distributed._symmetric_memory as _sm  # noqa: F401
            return hasattr(_sm, "rendezvous") and hasattr(_sm, "empty")
        except Exception:
            return False

    def sync(self, flat: torch.Tensor, accumulators=()):
        """flat (fp32, contiguous, CUDA) <- mean over ranks, in place; accumulators' sample counts are scaled."""
        assert flat.dtype == torch.float32 and flat.is_contiguous() and flat.numel() <= self.max_floats
        abi = self._abi
        abi.check(abi.lib.ecf_norm_exchange_p2p(flat.data_ptr(), flat.numel(), self._ptrs, self.max_floats, self.rank, self.world,
                                                torch.cuda.current_stream(flat.device).cuda_stream))
        for a in accumulators:
            a.nsamples *= self.world


def sync_block_hessians(accumulators, group=None):
    accs = list(accumulators)
    # accumulators fed by identical inputs share one H tensor (HessianBatch): reduce every distinct matrix once
    uniq, owner = [], {}
    for a in accs:
        key = a.H.data_ptr()
        if key not in owner:
            owner[key] = len(uniq)
            uniq.append(a)
    totals = allreduce_running_means([a.H for a in uniq], [a.nsamples for a in uniq], group)
    for a in accs:
        a.nsamples = totals[owner[a.H.data_ptr()]]


def row_sharded_select(W: torch.Tensor, select_rows: Callable[[torch.Tensor], None], group=None):
    """Prune this rank's row shard of W in place with ``select_rows(shard)`` and all-gather the shards so
    every rank ends up with the fully pruned matrix.  Falls back to replicated selection when the rows do not
    divide evenly."""
    rank, world = rank_world(group)
    R = W.shape[0]
    if world == 1 or R % world != 0 or not W.is_contiguous():
        select_rows(W)
        return W
    r0, r1 = row_range(R, rank, world)
    shard = W[r0:r1]
    select_rows(shard)
    dist.all_gather_into_tensor(W, shard.clone(), group=group)
    return W
