"""Host rows of the path checked on CPU against fixtures from the unmodified reference (tests/gen_golden.py):
the product's allocator (A15), return_sparsity grouping (A14/A16), the NormBatch in-place guard, block-output
indexing of the sweep, and the UPop positional-argument quirk."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from ecoflap_b200.layer_sparsity import LayerSparsity, _UniformSparsity


def _ls():
    return LayerSparsity.__new__(LayerSparsity)


def test_product_allocator_bit_exact_vs_reference_fixture(golden):
    """LayerSparsity.compute_the_sparsity_per_group (layer_single_base_pruner.py:247-314) -- same torch CPU ops, same
    promotions, so the ratios must be the reference's to the last bit (12 cases, 4 at BLIP-2 scale: 87 groups)."""
    g = golden("allocator")
    ls = _ls()
    for name in [str(c) for c in g["cases"]]:
        scores = {f"g{i}": torch.tensor(float(v), dtype=torch.float32) for i, v in enumerate(g[f"{name}__scores"])}
        sizes = {f"g{i}": int(v) for i, v in enumerate(g[f"{name}__sizes"])}
        res = ls.compute_the_sparsity_per_group(int(g[f"{name}__keep"]), scores, sizes,
                                                max_sparsity_per_layer=float(g[f"{name}__maxsp"]))
        got = np.array([res[k] for k in sizes], dtype=np.float64)
        assert np.array_equal(got, g[f"{name}__res"]), (name, np.abs(got - g[f"{name}__res"]).max())


def test_product_allocator_overshoot_toy():
    """SURVEY A15 known answer: scores {1,3}, sizes {100,200}, keep 150, max 0.6 -> the '+=' overshoot branch (:301)
    leaves 152 parameters: ratios 0.52 / 0.48 (as fp32-valued floats)."""
    res = _ls().compute_the_sparsity_per_group(150, {"a": torch.tensor(1.0), "b": torch.tensor(3.0)}, {"a": 100, "b": 200},
                                               max_sparsity_per_layer=0.6)
    assert res["a"] == float(np.float32(1) - np.float32(48) / np.float32(100))
    assert res["b"] == float(np.float32(1) - np.float32(104) / np.float32(200))
    kept = (1 - res["a"]) * 100 + (1 - res["b"]) * 200
    assert round(kept) == 152


def test_return_sparsity_matches_reference_fixture(golden):
    """return_sparsity with a pre-filled importance_measure (the fixture's scores): grouping + aggregation + allocation."""
    g = golden("return_sparsity")
    for method in [str(c) for c in g["cases"]]:
        keys = [str(k) for k in g[f"{method}__keys"]]
        groups = [str(k) for k in g[f"{method}__groups"]]

        class Toy(nn.Module):
            def __init__(self):
                super().__init__()
                self.blocks = nn.ModuleList(
                    [nn.ModuleDict({"a": nn.Linear(24, 40, bias=False), "b": nn.Linear(40, 24, bias=False)}) for _ in range(5)])

        m = Toy()
        with torch.no_grad():
            for k, p in m.named_parameters():
                p.copy_(torch.from_numpy(g[f"{method}__W__{k}"]))
        mapping = dict(zip(keys, groups))
        ls = LayerSparsity(m, None, None, 8, float(g["sparsity"]), float(g["maxsp"]), method, 1, 1e-3, mapping)
        ls.importance_measure = {k: torch.FloatTensor([v]) for k, v in zip(keys, g[f"{method}__impsum"])}
        res = ls.return_sparsity()
        got = np.array([res[k] for k in keys])
        # the fixture's impsum is the fp32 sum of the reference's per-element score tensor; feeding it back as a 1-element
        # tensor is what the product does (sum(|W| g) == g sum|W|), so the allocation must agree to fp32 rounding
        np.testing.assert_allclose(got, g[f"{method}__res"], rtol=0, atol=2e-6, err_msg=method)


def test_empty_mapping_is_uniform():
    ls = LayerSparsity(nn.Linear(4, 4), None, None, 8, 0.5, 0.6, "MEZO-GradOnly_sum", layer_to_group_mapping={})
    res = ls.return_sparsity()
    assert isinstance(res, _UniformSparsity) and res["anything"] == 0.5


def test_norm_batch_detects_in_place_update_of_a_hooked_input():
    """the hook passes inp[0].detach() (shares the version counter), so an in-place update between the
    hook and the deferred launch raises instead of silently corrupting the norms."""
    from ecoflap_b200.accumulators import NormBatch

    a = torch.randn(2, 3, 8)
    nb = NormBatch()
    nb.add(a.detach().reshape(-1, 8), torch.zeros(8), 0.0, 0.5)
    a.add_(1)
    with pytest.raises(RuntimeError, match="modified in place"):
        nb.flush()
    # ...whereas `.data` would not have noticed (its version counter is its own): the regression this test guards
    b = torch.randn(2, 3, 8)
    v = b.data._version
    b.add_(1)
    assert b.data._version == v and b.detach()._version != v


def test_run_block_only_indexes_tuple_outputs():
    """transformers >= 5 LlamaDecoderLayer returns a bare tensor; out[0] would drop the batch dimension."""
    from ecoflap_b200.pruners.sweep import SweepSpec, _run_block

    spec = SweepSpec(select="row", block_output_index=0)
    x = torch.randn(3, 5, 8)
    assert _run_block(lambda inp: inp * 2, x, {}, spec).shape == (3, 5, 8)
    assert _run_block(lambda inp: (inp * 2, None), x, {}, spec).shape == (3, 5, 8)


def test_upop_quirk_is_the_default():
    """UPop/pruners/wanda_pruner.py:707-717 passes `task` and the mapping positionally into num_noise / noise_eps, so
    the reference's BLIPBertLayerWandaPruner always allocates uniformly; the drop-in default must do the same."""
    import inspect

    from ecoflap_b200.pruners.upop import BLIPBertLayerWandaPruner

    assert inspect.signature(BLIPBertLayerWandaPruner.__init__).parameters["reference_uniform_quirk"].default is True


def test_zeroth_order_prefix_cache_protocol_matches_full_forwards_on_cpu():
    """The block-prefix cache of the zeroth-order loop (layer_sparsity._ReplayedLoss) without a GPU: the cut variants are
    run eagerly instead of being replayed from graphs, with the same bookkeeping as __call__ (static batch / cache buffers,
    write-back on a batch switch, dirty-from tracking), for every scored layer and three batches of the toy BLIP-2,
    including the inexact restore that leaves a rounding residue in the weights.  Every loss must equal the full forward's."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import e2e_cases as cases
    from ecoflap_b200.layer_sparsity import _ReplayedLoss, _copy_struct, _map_struct

    torch.manual_seed(0)
    m = cases.blip2_model().eval()
    batches = list(cases.blip2_loader())[:3]

    def loss_func(model, b, cuda):
        return model(b)["loss"], len(b["image"])

    names = [n for n, p in m.named_parameters() if n.endswith("weight") and ("blocks." in n or ".block." in n) and p.dim() == 2]
    r = _ReplayedLoss(loss_func, m, "cpu", names)
    r.prefix, r.static, r._stride = True, True, 2
    r._find_blocks(names)
    params = dict(m.named_parameters())
    with torch.no_grad():
        r._order_blocks(batches[0])
        assert r._blocks is not None and len(r._blocks) == 7 and r._cuts[0] == 0 and len(r._cuts) >= 4
        assert [type(b).__name__ for b in r._blocks] == ["EvaBlock"] * 3 + ["T5Block"] * 4  # T5Blocks, not their sub-layers
        for b in batches:
            k = id(b)
            r._run(r._cache.setdefault(k, {}), 0, b)
            r._batches[k], r._dirty[k] = (b, b), 0
            if r._sbatch is None:
                r._sbatch = _map_struct(b, lambda t: t.clone())
                r._scache = {j: _map_struct(v, lambda t: t.clone()) for j, v in r._cache[k].items()}
                r._cur, r._newer, r._loaded = k, set(), set(r._scache)
            assert r._same_struct(r._sbatch, b)

        def evaluate(b, name):
            k, blk = id(b), r._block_of[name]
            cut = max(c for c in r._cuts if c <= min(blk, r._dirty[k]))
            if r._cur != k:
                for j in r._newer:
                    _copy_struct(r._cache[r._cur][j], r._scache[j])
                _copy_struct(r._sbatch, b)
                r._cur, r._newer, r._loaded = k, set(), set()
            for src in set(r._ret[cut].values()):
                if src not in r._loaded and src not in r._newer:
                    _copy_struct(r._scache[src], r._cache[k][src])
                    r._loaded.add(src)
            loss, _ = r._run(r._scache, cut, r._sbatch)
            upd = {j for j in r._boundary if j not in r._ret[cut]}
            r._newer |= upd
            r._loaded -= upd
            r._dirty[k] = blk
            return float(loss), cut

        cuts_used = set()
        for name in names:
            p = params[name]
            for b in batches:
                z = torch.randn_like(p) * 1e-2
                p.data.add_(z)
                got, cut = evaluate(b, name)
                assert got == float(loss_func(m, b, False)[0]), (name, cut)
                p.data.sub_(2 * z)
                got, cut = evaluate(b, name)
                assert got == float(loss_func(m, b, False)[0]), (name, cut)
                p.data.add_(z * 1.001)  # inexact restore
                cuts_used.add(cut)
        assert len(cuts_used) >= 3 and max(cuts_used) > 0  # prefixes were really skipped


def test_block_replay_eligibility_rules():
    """N2 (pruners/sweep.py): which calibration sweeps are replayed from a graph."""
    from ecoflap_b200.pruners.sweep import SweepSpec, _BlockReplay

    spec = SweepSpec(select="row")
    x = [torch.zeros(1, 5, 8) for _ in range(32)]
    caches = [{"mask": None} for _ in range(32)]
    ok = _BlockReplay.eligible
    assert not ok(x, caches, list(range(32)), spec, 16)  # CPU tensors: never
    if torch.cuda.is_available():
        xg = [t.cuda() for t in x]
        assert ok(xg, caches, list(range(32)), spec, 16)
        assert not ok(xg, caches, list(range(16)), spec, 16)       # fewer than two groups
        assert not ok(xg, caches, list(range(31)), spec, 16)       # the group size must divide the batch count
        rot = [{"position_embeddings": (torch.ones(3), torch.ones(3))} for _ in range(32)]
        assert not ok(xg, rot, list(range(32)), spec, 16)          # per-sample tuples of tensors: eager
        shared = (torch.ones(3), torch.ones(3))
        assert ok(xg, [{"position_embeddings": shared} for _ in range(32)], list(range(32)), spec, 16)  # one shared object
