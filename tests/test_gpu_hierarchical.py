"""N1: hierarchical two-step embedding ≙ Embedder::from_hkgraph / h_embed (embedder.rs:120-133,194-295).
The projection-init kernel is checked against the formula of embedder.rs:245-269 (noise statistics, clipping, the
first-step rows copied verbatim) and the whole two-step path is run end to end."""
import numpy as np
import pytest

import annembed_b200 as A
import workloads

pytestmark = pytest.mark.gpu


def make_projection(n=6000, n_small=600, k=8, seed=0):
    x, _ = workloads.gaussian_mixture(n, 50, n_clusters=6, seed=seed, sub_dim=10, intrinsic=4, spread=20.0, sigma=1.0,
                                      lo=-1e4, hi=1e4)
    idx, dist = workloads.knn_exact(x, k, device="cuda")
    large = A.KGraph.from_knn(idx, dist)
    idx_s, dist_s = workloads.knn_exact(x[:n_small], k, device="cuda")          # upper-layer points are indexed first
    small = A.KGraph.from_knn(idx_s, dist_s)
    # projection of every node on the small graph: nearest of the first n_small points
    import torch
    xt = torch.as_tensor(x, device="cuda")
    d = torch.cdist(xt, xt[:n_small])
    pd, pn = d.min(dim=1)
    return x, A.KGraphProjection(small, large, pn.cpu().numpy(), pd.cpu().numpy())


def test_projection_init_follows_the_reference_formula():
    x, proj = make_projection()
    n, ns, d = proj.large_graph.get_nb_nodes(), proj.small_graph.get_nb_nodes(), 3
    first = np.random.default_rng(1).uniform(-3, 3, size=(ns, d)).astype(np.float32)
    ctx = A.CudaContext(A.EmbedderParams(asked_dim=d, seed=21))
    ctx.set_graph_csr(*proj.large_graph.get_neighbours())
    med = proj.get_projection_distance_median()
    ctx.set_embedding_from_projection(first, proj.proj_node, proj.proj_dist, med)
    y = ctx.get_embedding()
    np.testing.assert_array_equal(y[:ns], first)                                # embedder.rs:249-253
    delta = y[ns:] - first[proj.proj_node[ns:].astype(np.int64)]
    corr = np.sqrt(proj.proj_dist[ns:] / med / d)[:, None]                       # :262-263
    assert np.abs(delta).max() <= 2.0 + 1e-6                                     # clip(.., 2.)
    with np.errstate(divide="ignore", invalid="ignore"):
        z = delta / corr                                                         # ~ N(0,1) where not clipped
    unclipped = (np.abs(delta) < 2.0 - 1e-6) & (corr > 0)
    assert unclipped.mean() > 0.9
    assert abs(z[unclipped].mean()) < 0.03 and abs(z[unclipped].std() - 1.0) < 0.05
    ok = unclipped.all(axis=1)
    assert abs(np.corrcoef(z[ok, 0], z[ok, 1])[0, 1]) < 0.05                    # coordinates independent
    # deterministic for a seed, different for another
    ctx.set_embedding_from_projection(first, proj.proj_node, proj.proj_dist, med)
    np.testing.assert_array_equal(ctx.get_embedding(), y)
    with pytest.raises(A.AnnembedCudaError):
        bad = proj.proj_node.copy(); bad[-1] = ns
        ctx.set_embedding_from_projection(first, bad, proj.proj_dist, med)


def test_h_embed_end_to_end():
    x, proj = make_projection()
    ns = proj.small_graph.get_nb_nodes()
    params = A.EmbedderParams(nb_grad_batch=8, grad_factor=3, grad_step=1.0, dmap_init=False)
    emb = A.Embedder.from_hkgraph(proj, params, initial_embedding=workloads.pca_init(x[:ns], 2))
    assert emb.embed() == 1
    y = emb.get_embedded()
    assert y.shape == (6000, 2) and np.isfinite(y).all()
    # first step ran grad_factor * nb_grad_batch batches on the small graph (embedder.rs:203-205)
    per_batch = 10 * len(proj.small_graph.col)
    assert abs(emb.first_step_stats["positive_samples"] / (24 * per_batch) - 1) < 0.02
    assert abs(emb.stats["positive_samples"] / (8 * 10 * len(proj.large_graph.col)) - 1) < 0.02
    # neighbours of the large graph are close in the layout
    rp, col, _ = proj.large_graph.get_neighbours()
    src = np.repeat(np.arange(6000), 8)
    d_nb = np.linalg.norm(y[src] - y[col.astype(np.int64)], axis=1).mean()
    rnd = np.random.default_rng(0).integers(0, 6000, size=(48000, 2))
    assert d_nb < 0.2 * np.linalg.norm(y[rnd[:, 0]] - y[rnd[:, 1]], axis=1).mean()
