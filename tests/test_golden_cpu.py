"""The committed golden vectors are what the oracle produces today (guards the fixture against drift)."""
import os

import numpy as np

from oracle import oracle

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hotpath_small.npz"))


def test_oracle_reproduces_golden_fixture():
    row_ptr, col, dist = G["row_ptr"], G["col"], G["dist"]
    for tag in "abc":
        rho, beta = G[f"w_{tag}_params"]
        s, p = oracle.edge_weights(row_ptr, col, dist, rho, beta)
        np.testing.assert_array_equal(s, G[f"w_{tag}_scale"])
        np.testing.assert_allclose(p, G[f"w_{tag}_p"], rtol=1e-6)
    es = oracle.embedded_scales(G["w_a_scale"])
    np.testing.assert_array_equal(es, G["emb_scale"])
    p = G["w_a_p"]
    for d in (2, 3, 15):
        for b in (1.0, 0.5):
            out = oracle.step_fixed(row_ptr, col, p, es, G[f"y0_d{d}"], b, 1.0, G[f"edges_d{d}"], G[f"negs_d{d}"])
            np.testing.assert_allclose(out, G[f"step_d{d}_b{b}_g1.0"], rtol=1e-6, atol=1e-7)
            ce = oracle.cross_entropy(row_ptr, col, p, es, G[f"y0_d{d}"], b)
            np.testing.assert_allclose(ce, float(G[f"ce_d{d}_b{b}"]), rtol=1e-12)


def test_dmap_oracle_reproduces_golden_fixture():
    from oracle import dmap
    D = np.load(os.path.join(os.path.dirname(__file__), "golden", "dmap_small.npz"))
    row_ptr, col, dist = G["row_ptr"], G["col"], G["dist"]
    vd, v, sw, normed = dmap.sym_kernel(row_ptr, col, dist, 12)
    np.testing.assert_allclose(vd, D["diag"], rtol=1e-6)
    np.testing.assert_allclose(v, D["val"], rtol=1e-6)
    np.testing.assert_allclose(sw, D["sw"], rtol=1e-6)
    np.testing.assert_array_equal(normed, D["normed"])
    y, lam, _ = dmap.dmap_layout_randomized(row_ptr, col, dist, asked_dim=2, omega=D["omega"])
    np.testing.assert_allclose(lam, D["sigma"], atol=1e-9)
    for c in range(2):
        s = np.sign(np.dot(y[:, c], D["layout_d2"][:, c]))
        np.testing.assert_allclose(s * y[:, c], D["layout_d2"][:, c], atol=1e-4)
