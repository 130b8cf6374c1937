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


def _zo_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ecoflap_b200 import dist as edist
    from ecoflap_b200.layer_sparsity import LayerSparsity

    class CpuLayerSparsity(LayerSparsity):
        """Host logic under test; the CUDA kernels are replaced by the reference's own expressions."""

        def zo_perturb_parameters(self, params, random_seed=1, scaling_factor=1, zo_eps=1e-3):
            torch.manual_seed(random_seed)  # layer_single_base_pruner.py:473-486
            for param in params:
                z = torch.normal(mean=0, std=1, size=param.data.size(), device=param.data.device, dtype=param.data.dtype)
                param.data = param.data + scaling_factor * z * zo_eps

        def _magnitude_sums(self, params):
            return [float(p.data.abs().sum()) for p in params], [float((p.data ** 2).sum()) for p in params]

        def _grad_accumulate(self, G, grads, square):  # layer_single_base_pruner.py:447-450
            for g_acc, gr in zip(G, grads):
                g_acc.add_(gr.detach().float() ** 2 if square else gr.detach().float().abs())

        def _score_sums(self, params, G, n_batches, mode):  # :463-469, summed
            out = []
            for p_, g_ in zip(params, G):
                gbar, w = g_ / n_batches, p_.data.float()
                s_ = {"grad_mag_sq": w ** 2 * gbar, "grad_mag_abs": w.abs() * gbar.abs(), "grad_only": gbar.abs()}[mode]
                out.append(s_.double().sum())
            return torch.stack(out)

    def make_model():
        torch.manual_seed(0)
        return torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(),
                                   torch.nn.Linear(16, 8), torch.nn.Tanh(), torch.nn.Linear(8, 4))

    g = torch.Generator().manual_seed(3)
    loader = [{"x": torch.randn(5, 12, generator=g), "y": torch.randn(5, 4, generator=g)} for _ in range(4)]

    def loss_func(model, batch, cuda_enabled):
        return ((model(batch["x"]) - batch["y"]) ** 2).mean(), batch["x"].shape[0]

    def run(sharded):
        model = make_model()
        mapping = {k: k for k, v in model.named_parameters() if v.dim() == 2}
        ls = CpuLayerSparsity(model, loader, loss_func, num_samples=12, original_sparsity=0.5, max_sparsity_per_layer=0.8,
                              score_method="MEZO-GradMagAbs_sum", num_noise=2, noise_eps=1e-3, layer_to_group_mapping=mapping)
        np.random.seed(42)
        real = edist.is_dist
        if not sharded:
            edist.is_dist = lambda: False
        try:
            scores = ls.compute_importance_scores_mezo(mapping)
        finally:
            edist.is_dist = real
        return scores, [p.data.clone() for p in model.parameters()], np.random.randint(1 << 30)

    s_seq, w_seq, next_seq = run(False)   # the reference's sequential loop, on every rank
    s_sh, w_sh, next_sh = run(True)       # layers l = rank (mod 2), merged by one all-reduce
    assert next_seq == next_sh            # the numpy seed stream advanced identically
    for a, b in zip(w_seq, w_sh):         # stage 2 sees the very same (inexactly restored) weights
        assert torch.equal(a, b)
    assert set(s_seq) == set(s_sh) and len(s_sh) == 4
    for k in s_seq:
        np.testing.assert_allclose(s_sh[k].numpy(), s_seq[k].numpy(), rtol=1e-3)
    assert edist.zo_draws_per_layer([5, 5, 5, 5], 12, 2) == 3 and edist.zo_draws_per_layer([8] * 16, 32, 1) == 4

    # first-order scores (A13): data parallel over the batches, one scalar per layer exchanged
    def run_first_order(sharded, method):
        model = make_model()
        mapping = {k: k for k, v in model.named_parameters() if v.dim() == 2}
        ls = CpuLayerSparsity(model, loader, loss_func, num_samples=15, original_sparsity=0.5, max_sparsity_per_layer=0.8,
                              score_method=method, layer_to_group_mapping=mapping)
        real = edist.is_dist
        if not sharded:
            edist.is_dist = lambda: False
        try:
            return ls.compute_importance_scores(mapping)
        finally:
            edist.is_dist = real

    for method in ("GradMagAbs_sum", "GradMagSquare_avg", "GradOnly_sum"):
        f_seq, f_sh = run_first_order(False, method), run_first_order(True, method)
        for k in f_seq:
            np.testing.assert_allclose(f_sh[k].numpy(), f_seq[k].numpy(), rtol=1e-5)
    dist.barrier()
    with open(os.path.join(out_dir, f"zo_ok_{rank}"), "w") as fh:
        fh.write("ok")
    dist.destroy_process_group()


def test_zeroth_order_loop_sharded_over_layers_world2(tmp_path):
    """the zeroth-order (layer, batch) grid sharded over ranks reproduces the sequential loop's seed stream
    and final weights exactly and its g-hat within tolerance (gloo, CPU; the CUDA kernels are replaced by the
    reference's expressions)."""
    port = 29650 + os.getpid() % 200
    mp.spawn(_zo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"zo_ok_{r}")) for r in range(2))
