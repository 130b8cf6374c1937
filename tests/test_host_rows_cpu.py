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
