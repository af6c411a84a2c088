"""CPU checks of the diffusion-map restatement (oracle/dmap.py): the invariants the reference itself
asserts or relies on, on a point cloud whose diffusion coordinates are known in closed form."""
import numpy as np
import pytest

import annembed_b200 as A
import workloads
from oracle import dmap


def strip_graph(n=3000, k=10, seed=0):
    """uniform points on a 4 x 1 strip: the first diffusion coordinates are cos(pi x / 4), cos(2 pi x / 4), ..."""
    rng = np.random.default_rng(seed)
    x = np.stack([rng.uniform(0, 4, n), rng.uniform(0, 1, n)], 1).astype(np.float32)
    idx, dist = workloads.knn_exact(x, k)
    g = A.KGraph.from_knn(idx, dist)
    return x, g


def test_kernel_normalisation_and_symmetry():
    x, g = strip_graph()
    row_ptr, col, dist = g.get_neighbours()
    vd, v, sw, normed = dmap.sym_kernel(row_ptr, col, dist)
    S = dmap.kernel_matrix(row_ptr, col, vd, v)
    assert abs(S - S.T).max() == 0
    # D^-1/2 K D^-1/2 has the eigenvector sqrt(degrees) with eigenvalue 1: the reference's own debug check,
    # diffmaps.rs:482-490 ("bad normalization" if off by more than 1e-3)
    np.testing.assert_allclose(S @ sw.astype(np.float64), sw, rtol=2e-5)
    assert abs(normed.mean() - 1) < 1e-5 and normed.min() > 0
    assert (v > 0).all() and (vd > 0).all()


def test_symmetrisation_follows_the_reference_triplets():
    """diffmaps.rs:522-539: every directed entry is pushed in both directions with max(w_ij, w_ji): mutual pairs end up
    with twice the maximum, one-way edges with their weight, the self edge with twice its weight."""
    row_ptr = np.array([0, 2, 3, 4], np.uint64)
    col = np.array([1, 2, 0, 1], np.uint32)                # 0->1, 0->2, 1->0, 2->1
    w = np.array([0.5, 0.25, 0.75, 0.125], np.float32)
    sym = dmap.symmetrise(row_ptr, col, w)
    np.testing.assert_array_equal(sym, np.array([0.75, 0.25, 0.75, 0.125], np.float32))
    q = dmap.sym_rowsum(row_ptr, col, sym, np.ones(3, np.float32))
    # node 0: (0,1)+(1,0) mutual -> 2*0.75, (0,2) one way 0.25, self 2  ; node 1: 1.5 + 0.125 + 2 ; node 2: 0.25 + 0.125 + 2
    np.testing.assert_allclose(q, [1.5 + 0.25 + 2, 1.5 + 0.125 + 2, 0.25 + 0.125 + 2])


def test_all_equal_rows_and_zero_scales():
    """diffmaps.rs:614-646 (rows whose distances are all equal or all zero get uniform weights 1/(k+1), self included)
    and :790-797 (zero scales take the mean)."""
    row_ptr = np.array([0, 2, 4, 6], np.uint64)
    col = np.array([1, 2, 0, 2, 0, 1], np.uint32)
    dist = np.array([0.0, 0.0, 0.5, 0.5, 0.2, 0.9], np.float32)
    s, mean = dmap.local_scales(row_ptr, dist)
    assert s[0] == mean and s[1] == np.float32(0.5)
    ws, w = dmap.kernel_weights(row_ptr, col, dist, s)
    np.testing.assert_allclose(ws, [1 / 3, 1 / 3, 1.0])
    np.testing.assert_allclose(w[:4], 1 / 3)
    assert w[4] > w[5] >= dmap.PROBA_MIN


def test_layout_recovers_the_strip_coordinates():
    x, g = strip_graph()
    y, lam, U = dmap.dmap_layout(*g.get_neighbours(), asked_dim=2)
    assert abs(lam[0] - 1) < 1e-5 and np.all(np.diff(lam) <= 1e-12)          # spectrum decreasing from 1 (diffmaps.rs:1172)
    assert abs(np.corrcoef(y[:, 0], np.cos(np.pi * x[:, 0] / 4))[0, 1]) > 0.97
    assert abs(np.corrcoef(y[:, 1], np.cos(2 * np.pi * x[:, 0] / 4))[0, 1]) > 0.95
    # set_data_box(., 10): centred, largest coordinate 5 (embedder.rs:1376-1408)
    np.testing.assert_allclose(y.mean(axis=0), 0, atol=1e-5)
    assert abs(np.abs(y).max() - 5) < 1e-5


def test_randomized_svd_approaches_the_exact_spectrum():
    x, g = strip_graph()
    row_ptr, col, dist = g.get_neighbours()
    vd, v, sw, normed = dmap.sym_kernel(row_ptr, col, dist)
    S = dmap.kernel_matrix(row_ptr, col, vd, v)
    lam = np.sort(np.abs(np.linalg.eigvalsh(S.toarray())))[::-1]
    s, U = dmap.subspace_svd(S, seed=3)
    # singular values of Q^T S interlace those of S from below; with the reference's settings (rank 20, 5 iterations)
    # and a kernel spectrum this flat (lambda_21 / lambda_1 = 0.985) they are only rough: that is the reference's init
    assert np.all(s <= lam[:20] + 1e-9) and s[0] > 0.95 and np.all(np.diff(s) <= 0)
    np.testing.assert_allclose(U.T @ U, np.eye(20), atol=1e-8)
    # more iterations converge to the exact spectrum (the restatement iterates the right operator)
    s40, _ = dmap.subspace_svd(S, nbiter=400, seed=3)
    np.testing.assert_allclose(s40[:3], lam[:3], atol=2e-4)
