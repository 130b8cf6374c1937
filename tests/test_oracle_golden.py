"""Pins the numpy oracle (oracle/ecoflap_oracle.py) against fixtures produced by the UNMODIFIED
reference classes (tests/gen_golden.py).  CPU only."""
import numpy as np
import pytest

import ecoflap_oracle as orc


def _cases(g):
    return [str(c) for c in g["cases"]]


def test_norm_accumulator_matches_reference(golden):
    g = golden("norm_accum")
    for name in _cases(g):
        nb = int(g[f"{name}__nb"])
        acc = orc.NormAccumulator(g[f"{name}__x0"].shape[-1])
        for i in range(nb):
            acc.add_batch(g[f"{name}__x{i}"])
            ref = g[f"{name}__s{i}"]
            assert acc.nsamples == int(g[f"{name}__n{i}"])
            # tolerance from north_star: norms within 1e-3 relative (summation order differs)
            np.testing.assert_allclose(acc.scaler_row, ref, rtol=1e-5, atol=1e-30)


def test_norm_final_equals_mean_sum_of_squares(golden):
    g = golden("norm_accum")
    for name in _cases(g):
        nb = int(g[f"{name}__nb"])
        tot, n = 0.0, 0
        for i in range(nb):
            x = g[f"{name}__x{i}"]
            tot = tot + orc.sqnorm_columns(x)
            n += 1 if x.ndim == 2 else x.shape[0]
        np.testing.assert_allclose(g[f"{name}__s{nb - 1}"], tot / n, rtol=1e-5, atol=1e-30)


def test_hessian_accumulator_matches_reference(golden):
    g = golden("hessian_accum")
    for name in _cases(g):
        nb = int(g[f"{name}__nb"])
        acc = orc.HessianAccumulator(g[f"{name}__x0"].shape[-1])
        xs = []
        for i in range(nb):
            xs.append(g[f"{name}__x{i}"])
            acc.add_batch(xs[-1])
            ref = g[f"{name}__H{i}"]
            scale = np.abs(ref).max()
            assert np.abs(acc.H - ref).max() <= 1e-5 * scale
        exact = orc.hessian_exact(xs)
        assert np.abs(exact - ref).max() <= 1e-5 * np.abs(ref).max()


def test_zo_perturb_bit_exact(golden):
    g = golden("zo_perturb")
    eps = float(g["eps"])
    for dt in _cases(g):
        w = g[f"{dt}__W0"]
        z = g[f"{dt}__z"]
        for step, sc in enumerate((1, -2, 1)):
            w = orc.zo_perturb(w, z, sc, eps, dt)
            ref = g[f"{dt}__W{step + 1}"]
            assert np.array_equal(w.view(np.uint32), ref.view(np.uint32)), (dt, step)
        # the reference's restore is inexact in low precision: keep that visible
        if dt != "fp32":
            assert not np.array_equal(w, g[f"{dt}__W0"])


def test_allocator_matches_reference(golden):
    g = golden("allocator")
    for name in _cases(g):
        scores = {f"g{i}": v for i, v in enumerate(g[f"{name}__scores"])}
        sizes = {f"g{i}": int(v) for i, v in enumerate(g[f"{name}__sizes"])}
        res = orc.sparsity_per_group(int(g[f"{name}__keep"]), scores, sizes, float(g[f"{name}__maxsp"]))
        got = np.array([res[k] for k in sizes])
        ref = g[f"{name}__res"]
        # The reference sums ~1e9-sized fp32 values with torch's CPU cascade sum, whose lane order
        # depends on the host's vector width; one fp32 ulp (128 params at BLIP-2 scale) of the
        # running total can add or skip a loop iteration, so ratios agree to ~1e-6 absolute.
        np.testing.assert_allclose(got, ref, rtol=0, atol=5e-6, err_msg=name)


def test_allocator_known_answer_toy(golden):
    # scores {1,3}, sizes {100,200}, keep 150, max 0.6 -> {0.52, 0.48}
    res = orc.sparsity_per_group(150, {"a": 1.0, "b": 3.0}, {"a": 100, "b": 200}, 0.6)
    assert res["a"] == pytest.approx(float(np.float32(0.52)), abs=1e-7)
    assert res["b"] == pytest.approx(float(np.float32(0.48)), abs=1e-7)
    kept = (1 - res["a"]) * 100 + (1 - res["b"]) * 200
    assert round(kept) == 152  # the reference's '+=' overshoot bug is reproduced, not fixed


def test_return_sparsity_matches_reference(golden):
    g = golden("return_sparsity")
    for method in _cases(g):
        keys = [str(k) for k in g[f"{method}__keys"]]
        groups = [str(k) for k in g[f"{method}__groups"]]
        mapping = dict(zip(keys, groups))
        numel = dict(zip(keys, (int(v) for v in g[f"{method}__numel"])))
        ghat = dict(zip(keys, g[f"{method}__ghat"]))
        comp, agg = method.split("_")
        sums = {}
        for k in keys:
            sa, sq = orc.abs_and_square_sums(g[f"{method}__W__{k}"])
            gk = np.float32(ghat[k])
            if comp == "MEZO-GradOnly":
                sums[k] = gk
            elif comp == "MEZO-GradMagAbs":
                sums[k] = np.float32(sa * float(gk))
            else:
                sums[k] = np.float32(sq * float(gk) ** 2)
        np.testing.assert_allclose(
            np.array([sums[k] for k in keys], dtype=np.float32), g[f"{method}__impsum"], rtol=1e-5
        )
        res = orc.layer_sparsity_from_scores(sums, numel, mapping, float(g["sparsity"]), float(g["maxsp"]), agg)
        np.testing.assert_allclose([res[k] for k in keys], g[f"{method}__res"], rtol=0, atol=5e-4, err_msg=method)


def test_obs_prune_matches_reference(golden):
    g = golden("obs_prune")
    for name in _cases(g):
        dt = str(g[f"{name}__dtype"])
        W, H, s = g[f"{name}__W"], g[f"{name}__H"], float(g[f"{name}__s"])
        pn, pm = (int(v) for v in g[f"{name}__nm"])
        Wout, mask = orc.obs_prune(W, H, s, out_dtype=dt, prune_n=pn, prune_m=pm)
        ref = g[f"{name}__Wout"]
        if pn:  # n:m: exactly n zeros in every group of m
            assert ((ref.reshape(ref.shape[0], -1, pm) == 0).sum(-1) >= pn).all(), name
        agree = ((Wout == 0) == (ref == 0)).mean()
        assert agree >= 0.995, (name, agree)
        rel = np.linalg.norm(Wout - ref) / np.linalg.norm(ref)
        assert rel <= 2e-2, (name, rel)  # LAPACK builds differ (numpy/OpenBLAS vs torch/MKL)
        assert abs((ref == 0).mean() - (Wout == 0).mean()) < 2e-3


def test_select_steps_match_reference_prune_loop(golden):
    """(W_before, scaler_row) captured inside the reference's own _prune loops (tests/gen_golden_e2e.py) ->
    the oracle's select must reproduce the reference's pruned weights bit for bit."""
    g = golden("e2e_pruners")
    for i in range(int(g["vit_wanda__nsteps"])):  # per-LAYER threshold, wanda_pruner.py:553-558
        W, s, ref = g[f"vit_wanda__step{i}__W"], g[f"vit_wanda__step{i}__s"], g[f"vit_wanda__step{i}__Wafter"]
        got, mask, _ = orc.wanda_prune_layer(W, s, 0.5)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), i
        assert mask.sum() >= int(W.size * 0.5) + 1
    for i in range(int(g["t5_wanda__nsteps"])):  # per-ROW stable sort, wanda_pruner.py:272-279
        W, s, ref = g[f"t5_wanda__step{i}__W"], g[f"t5_wanda__step{i}__s"], g[f"t5_wanda__step{i}__Wafter"]
        got, mask = orc.wanda_prune_rows(W, s, 0.5)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), i
        assert (mask.sum(axis=1) == int(W.shape[1] * 0.5)).all()


def test_nm_select_oracle_matches_the_reference_expression():
    """A6: the oracle's n:m mask against the reference's own expression (wanda_pruner.py:265-270:
    ``W_mask.scatter_(1, ii + torch.topk(tmp, prune_n, dim=1, largest=False)[1], True)`` per group of prune_m columns)
    evaluated with torch on tie-free scores, where torch.topk's unspecified tie order cannot matter."""
    import torch

    rng = np.random.default_rng(7)
    for (R, C, n, m) in [(16, 64, 2, 4), (9, 96, 4, 8), (5, 50, 2, 3), (4, 40, 1, 4), (3, 32, 5, 16)]:
        W = (rng.standard_normal((R, C)) * 0.02).astype(np.float32)
        s = (rng.random(C) + 0.1).astype(np.float32)
        M = orc.wanda_metric(W, s)
        assert len(np.unique(M)) == M.size  # tie-free
        W_metric = torch.abs(torch.from_numpy(W)) * torch.sqrt(torch.from_numpy(s).reshape((1, -1)))
        W_mask = (torch.zeros_like(W_metric) == 1)
        for ii in range(W_metric.shape[1]):
            if ii % m == 0:
                tmp = W_metric[:, ii:(ii + m)].float()
                W_mask.scatter_(1, ii + torch.topk(tmp, n, dim=1, largest=False)[1], True)
        assert np.array_equal(orc.nm_select_mask(M, n, m), W_mask.numpy()), (R, C, n, m)


def test_global_mask_oracle_matches_the_reference():
    """N3 (next row): the oracle's restatement of BLIPT5GlobalPruner.get_mask / get_layerwise_mask (global_pruner.py:
    116-157) against fixtures generated from the unmodified reference class (tests/gen_golden_global.py), including ties
    and protection of the top (1 - max_sparsity) fraction per layer."""
    g = np.load("tests/golden/global_mask.npz")
    names = [str(n) for n in g["names"]]
    for case in range(4):
        p, max_sp = float(g[f"c{case}__p"]), float(g[f"c{case}__max_sp"])
        scores = {n: g[f"c{case}__score{i}"] for i, n in enumerate(names)}
        masks, _ = orc.global_get_mask(scores, p, max_sp)
        lw = orc.global_layerwise_mask(scores, p)
        for i, n in enumerate(names):
            assert np.array_equal(masks[n], g[f"c{case}__mask{i}"]), (case, n)
            assert np.array_equal(lw[n], g[f"c{case}__lw{i}"]), (case, n)
    assert orc.global_iteration_ratios(0.5, 3) == [0.5 ** 3, 0.5 ** 1.5, 0.5]
