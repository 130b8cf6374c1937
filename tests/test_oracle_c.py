"""The plain-C (OpenMP) oracle against the numpy oracle and the reference-generated fixtures.  CPU only."""
import numpy as np
import pytest

import c_oracle
import ecoflap_oracle as orc


@pytest.mark.parametrize("dt", ["fp32", "fp16", "bf16"])
def test_c_row_and_layer_prune_match_numpy_oracle(dt):
    rng = np.random.default_rng(3)
    W = orc.round_to((rng.standard_normal((40, 96)) * 0.02).astype(np.float32), dt)
    W[:, 48:] = W[:, :48]  # ties
    s = (rng.random(96).astype(np.float32) + 0.1)
    s[5] = 0
    for sp in (0.5, 0.3):
        Ws = c_oracle.to_storage(W, dt)
        c_oracle.wanda_row_prune(Ws, dt, s, orc.row_k(96, sp))
        ref, _ = orc.wanda_prune_rows(W, s, sp)
        assert np.array_equal(c_oracle.from_storage(Ws, dt), ref)
        Ws = c_oracle.to_storage(W, dt)
        th = c_oracle.wanda_layer_prune(Ws, dt, s, orc.layer_kth_index(W.size, sp))
        ref, _, thres = orc.wanda_prune_layer(W, s, sp)
        assert th == float(thres)
        assert np.array_equal(c_oracle.from_storage(Ws, dt), ref)


def test_c_sqnorm_matches_reference_golden(golden):
    g = golden("norm_accum")
    for name in [str(c) for c in g["cases"]]:
        dt = str(g[f"{name}__dtype"])
        C = g[f"{name}__x0"].shape[-1]
        s = np.zeros(C, dtype=np.float32)
        n = 0
        for i in range(int(g[f"{name}__nb"])):
            x = g[f"{name}__x{i}"]
            b = 1 if x.ndim == 2 else x.shape[0]
            c_oracle.sqnorm_accum(c_oracle.to_storage(x.reshape(-1, C), dt), dt, s, n, b)
            n += b
            np.testing.assert_allclose(s, g[f"{name}__s{i}"], rtol=1e-5, atol=1e-30)
