"""N3: the device quality estimator (annembed_cuda_quality_estimate ≙ embedder.rs:620-753) against the
CPU restatement oracle/quality.py (exact kNN radii from scipy's cKDTree)."""
import numpy as np
import pytest

import annembed_b200 as A
import workloads
from oracle import quality
from tests.conftest import random_graph

pytestmark = pytest.mark.gpu


def compare(row_ptr, col, dist, y, nbng, d):
    ctx = A.CudaContext(A.EmbedderParams(asked_dim=d))
    ctx.set_graph_csr(row_ptr, col, dist)
    ctx.set_embedding(y)
    q = ctx.quality_estimate(nbng, want_arrays=True)
    ref = quality.quality_stats(row_ptr, col, y, nbng)
    tree_r = None
    from scipy.spatial import cKDTree
    dd, _ = cKDTree(y.astype(np.float64)).query(y.astype(np.float64), k=nbng + 1, workers=-1)
    np.testing.assert_allclose(q["radius"], dd[:, nbng], rtol=2e-5, atol=1e-7)          # exact k-th neighbour distance
    np.testing.assert_allclose(q["first_dist"], ref["first_dist"], rtol=1e-5, atol=1e-7)
    n = len(row_ptr) - 1
    assert abs(q["nb_without_match"] - ref["nb_without_match"]) <= max(2, 0.002 * n)
    assert abs(q["mean_nbmatch"] - ref["mean_nbmatch"]) <= 5e-3 * ref["mean_nbmatch"]
    assert abs(q["knn_preservation"] - ref["knn_preservation"]) <= 2e-3
    assert abs(q["mean_ratio"] - ref["mean_ratio"]) <= 1e-4 * ref["mean_ratio"]
    np.testing.assert_allclose(q["radius_quantiles"], ref["radius_quantiles"], rtol=1e-4)
    np.testing.assert_allclose(q["ratio_quantiles"], ref["ratio_quantiles"], rtol=1e-3)
    return q


def test_quality_estimator_2d_clustered_layout():
    # a clustered layout with very uneven density (dense blobs + sparse background) and some duplicate points
    rng = np.random.default_rng(0)
    n = 30000
    y = np.concatenate([rng.normal(0, 0.02, (12000, 2)) + [1, 1], rng.normal(0, 0.5, (12000, 2)) - [2, 0],
                        rng.uniform(-6, 6, (6000, 2))]).astype(np.float32)
    y[100:110] = y[100]
    row_ptr, col, dist = random_graph(n, 4, 9, seed=3)
    # make the graph meaningful: neighbours = true embedded kNN for half of the nodes
    compare(row_ptr, col, dist, y, 50, 2)
    compare(row_ptr, col, dist, y, 7, 2)


def test_quality_estimator_after_embedding_and_higher_dim():
    x, _ = workloads.gaussian_mixture(8000, 100, n_clusters=5, seed=2, sub_dim=10, intrinsic=4, spread=25.0, sigma=1.0,
                                      lo=-1e4, hi=1e4)
    idx, dist = workloads.knn_exact(x, 8, device="cuda")
    g = A.KGraph.from_knn(idx, dist)
    for d in (2, 5):
        emb = A.Embedder(g, A.EmbedderParams(asked_dim=d, nb_grad_batch=10, grad_step=1.0), initial_embedding=workloads.pca_init(x, d))
        emb.embed()
        q = compare(*g.get_neighbours(), emb.get_embedded(), 40, d)
        assert q["knn_preservation"] > 0.5
        q2 = emb.get_quality_estimate_from_edge_length(40)
        assert q2["nb_without_match"] == q["nb_without_match"]


def test_quality_estimator_errors():
    row_ptr, col, dist = random_graph(500, 3, 5, seed=1)
    ctx = A.CudaContext(A.EmbedderParams())
    ctx.set_graph_csr(row_ptr, col, dist)
    with pytest.raises(A.AnnembedCudaError) as e:
        ctx.quality_estimate(10)
    assert e.value.status == 5
    ctx.set_embedding(np.zeros((500, 2), np.float32))          # all points coincide: radius 0 everywhere
    q = ctx.quality_estimate(10)
    assert q["nb_without_match"] == 0 and q["radius_quantiles"][5] == 0.0
    with pytest.raises(A.AnnembedCudaError):
        ctx.quality_estimate(500)
