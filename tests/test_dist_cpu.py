"""World-size-2 gloo tests of the multi-GPU partitioning logic (ecoflap_b200/dist.py) on CPU tensors: batch-sharded
running means merged by one all-reduce, and row-sharded selection rebuilt by an all-gather.  The kernels are replaced
by numpy-oracle callables here; the same dist code drives the CUDA kernels under NCCL on the box."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ecoflap_oracle as orc
    from ecoflap_b200 import dist as edist

    rng = np.random.default_rng(0)  # same stream on every rank: the full calibration set
    C, n_batches, B = 96, 6, 4
    batches = [rng.standard_normal((B, 7, C)).astype(np.float32) for _ in range(n_batches)]
    W = (rng.standard_normal((8, C)) * 0.02).astype(np.float32)

    class Acc:  # WrappedGPT-like, numpy oracle inside
        def __init__(self):
            self.ref = orc.NormAccumulator(C)
            self.scaler_row = torch.zeros(C)
            self.nsamples = 0

        def add(self, x):
            self.scaler_row = torch.from_numpy(self.ref.add_batch(x).copy())
            self.nsamples = self.ref.nsamples

    accs = [Acc(), Acc()]
    mine = edist.shard_indices(n_batches, rank, world)
    assert mine == list(range(rank, n_batches, world))
    for j in mine:
        accs[0].add(batches[j])
        accs[1].add(2.0 * batches[j])
    edist.sync_block_norms(accs)
    full = orc.NormAccumulator(C)
    for x in batches:
        full.add_batch(x)
    assert accs[0].nsamples == n_batches * B == full.nsamples
    np.testing.assert_allclose(accs[0].scaler_row.numpy(), full.scaler_row, rtol=1e-5)
    np.testing.assert_allclose(accs[1].scaler_row.numpy(), 4.0 * full.scaler_row, rtol=1e-5)

    # packed variant: the block's accumulators are slices of one buffer, exchanged by ONE in-place all-reduce; with
    # known totals nothing is read back (capturable in a CUDA graph on the box)
    class PAcc:
        def __init__(self, scale):
            self.columns, self.dev, self.nsamples, self.scale = C, torch.device("cpu"), 0, scale
            self.scaler_row = torch.zeros(C)
            self.ref = orc.NormAccumulator(C)

        def add(self, x):  # in place: scaler_row is a view of the packed buffer
            self.scaler_row.copy_(torch.from_numpy(self.ref.add_batch(self.scale * x).copy()))
            self.nsamples = self.ref.nsamples

    paccs = [PAcc(1.0), PAcc(2.0)]
    flat = edist.pack_block_norms(paccs)
    assert flat.numel() == 2 * C and paccs[1].scaler_row.data_ptr() == flat.data_ptr() + 4 * C
    for j in mine:
        for a in paccs:
            a.add(batches[j])
    edist.sync_packed_norms(flat, paccs)
    assert paccs[0].nsamples == n_batches * B
    np.testing.assert_allclose(paccs[0].scaler_row.numpy(), full.scaler_row, rtol=1e-5)
    np.testing.assert_allclose(flat[C:].numpy(), 4.0 * full.scaler_row, rtol=1e-5)
    taccs = [Acc(), Acc()]
    for j in mine:
        taccs[0].add(batches[j])
        taccs[1].add(2.0 * batches[j])
    edist.sync_block_norms(taccs, totals=[n_batches * B] * 2)
    np.testing.assert_allclose(taccs[1].scaler_row.numpy(), 4.0 * full.scaler_row, rtol=1e-5)
    assert taccs[0].nsamples == n_batches * B

    # Hessians go through the same plumbing
    class HAcc:
        def __init__(self):
            self.ref = orc.HessianAccumulator(C)
            self.H = torch.zeros(C, C)
            self.nsamples = 0

        def add(self, x):
            self.ref.add_batch(x)
            self.H = torch.from_numpy(np.array(self.ref.H, dtype=np.float32))
            self.nsamples = self.ref.nsamples

    h = HAcc()
    for j in mine:
        h.add(batches[j])
    edist.sync_block_hessians([h])
    hfull = orc.HessianAccumulator(C)
    for x in batches:
        hfull.add_batch(x)
    scale = np.abs(hfull.H).max()
    assert np.abs(h.H.numpy() - hfull.H).max() <= 1e-5 * scale

    # row-sharded select + all-gather == the unsharded select, bit for bit
    s = full.scaler_row
    Wt = torch.from_numpy(W.copy())

    def select_rows(shard):
        pruned, _ = orc.wanda_prune_rows(shard.numpy(), s, 0.5)
        shard.copy_(torch.from_numpy(pruned))

    edist.row_sharded_select(Wt, select_rows)
    want, _ = orc.wanda_prune_rows(W, s, 0.5)
    assert np.array_equal(Wt.numpy(), want)
    r0, r1 = edist.row_range(8, rank, world)
    assert (r0, r1) == (rank * 4, rank * 4 + 4)
    # rows that do not divide evenly fall back to the replicated select
    W7 = torch.from_numpy(W[:7].copy())
    edist.row_sharded_select(W7, select_rows)
    assert np.array_equal(W7.numpy(), orc.wanda_prune_rows(W[:7], s, 0.5)[0])
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_single_process_helpers():
    for p in (ROOT,):
        if p not in sys.path:
            sys.path.insert(0, p)
    from ecoflap_b200 import dist as edist

    assert not edist.is_dist()
    assert edist.rank_world() == (0, 1)
    W = torch.arange(12.0).reshape(4, 3)
    edist.row_sharded_select(W, lambda sh: sh.mul_(0))
    assert not W.any()
