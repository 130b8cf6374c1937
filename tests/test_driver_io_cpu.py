"""N4 -- the driver-side artefacts (LAVIS/evaluate_blip.py:344-389,438-472) written and read back on CPU tensors."""
import os

import torch
import yaml


class _Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.t5_model = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.Linear(8, 4))
        self.visual_encoder = torch.nn.Sequential(torch.nn.Linear(6, 6))
        self.head = torch.nn.Linear(4, 2)


def test_save_and_reload_pruning_outputs(tmp_path):
    from ecoflap_b200 import driver_io as io

    torch.manual_seed(0)
    model = _Tiny()
    with torch.no_grad():
        model.t5_model[0].weight[:, ::2] = 0
        model.visual_encoder[0].weight[0] = 0
    sd = {"t5_model.0.weight": 0.52, "t5_model.1.weight": 0.48}
    out = io.save_pruning_outputs(model, "job7", sparsity_dict=sd, start_time=None, root=str(tmp_path))
    assert sorted(os.listdir(tmp_path)) == ["pruned_checkpoint", "sparsity_dict", "training_statistics"]
    assert set(yaml.safe_load(open(out["training_statistics"]))) == {"memory", "time"}
    assert io.load_sparsity_dict(out["sparsity_dict"]) == sd
    # the uniform module is not a dict: no sparsity file (evaluate_blip.py:450)
    out2 = io.save_pruning_outputs(model, "job8", sparsity_dict=object(), root=str(tmp_path))
    assert "sparsity_dict" not in out2
    fresh = _Tiny()
    io.load_t5_pruned_checkpoint(fresh, out["checkpoint"])
    io.load_vit_pruned_checkpoint(fresh, out["checkpoint"])
    assert torch.equal(fresh.t5_model[0].weight, model.t5_model[0].weight)
    assert torch.equal(fresh.visual_encoder[0].weight, model.visual_encoder[0].weight)
    assert not torch.equal(fresh.head.weight, model.head.weight)  # only the two towers are re-loaded


def test_pack_unpack_sparse_roundtrip():
    from ecoflap_b200 import driver_io as io

    g = torch.Generator().manual_seed(1)
    for (R, C) in [(5, 16), (3, 50), (4, 7)]:
        W = torch.randn(R, C, generator=g).half()
        W[torch.rand(R, C, generator=g) < 0.5] = 0
        bits, vals = io.pack_sparse(W)
        assert bits.shape == (R, (C + 7) // 8) and vals.numel() == int((W != 0).sum())
        assert torch.equal(io.unpack_sparse(bits, vals, C), W)


def test_eva_clip_checkpoint_filter(tmp_path):
    """evaluate_eva_clip.py:414-423: only ``visual.*`` without ``blocks.39``."""
    from ecoflap_b200 import driver_io as io

    class Eva(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.visual = torch.nn.Module()
            self.visual.blocks = torch.nn.ModuleList([torch.nn.Linear(4, 4) for _ in range(41)])
            self.text = torch.nn.Linear(4, 4)

    m = Eva()
    kept = io.filter_eva_clip_checkpoint(m.state_dict())
    assert all(k.startswith("visual.") for k in kept) and not any("blocks.39" in k for k in kept)
    assert "visual.blocks.38.weight" in kept and "visual.blocks.40.weight" in kept and len(kept) == 2 * 40
    out = io.save_pruning_outputs(m, "eva", root=str(tmp_path), eva_clip=True)
    assert set(torch.load(out["checkpoint"])) == set(kept)
