"""GPU parity of the device diffusion-map initial layout against oracle/dmap.py.
The kernel construction is deterministic: <= 1e-5 relative like the edge weights.  The randomized SVD is compared with
the SAME Gaussian test matrix fed to both sides (annembed_cuda_dmap_set_test_matrix): singular values to 1e-4, layout
columns up to their sign."""
import numpy as np
import pytest

import annembed_b200 as A
import workloads
from oracle import dmap
from tests.conftest import random_graph
from tests.test_oracle_dmap import strip_graph

pytestmark = pytest.mark.gpu


def ctx_for(g, **kw):
    ctx = A.CudaContext(A.EmbedderParams(**kw))
    ctx.set_graph_csr(*g.get_neighbours())
    return ctx


def rel_err(a, b):
    return float(np.max(np.abs(a.astype(np.float64) - b) / np.maximum(np.abs(b), 1e-30)))


@pytest.mark.parametrize("case", ["strip", "ragged"])
def test_dmap_kernel_matches_oracle(case):
    if case == "strip":
        _, g = strip_graph(3000, 10)
        row_ptr, col, dist = g.get_neighbours()
    else:
        # ragged rows, rows with all-equal and all-zero distances (diffmaps.rs:614-646), one-way and mutual edges
        row_ptr, col, dist = random_graph(2500, 3, 14, seed=17)
        dist = dist.copy()
        for i in (5, 77, 300):
            dist[int(row_ptr[i]):int(row_ptr[i + 1])] = 0.0
        for i in (9, 1200):
            dist[int(row_ptr[i]):int(row_ptr[i + 1])] = 0.37
        g = A.KGraph(row_ptr, col, dist)
    ctx = ctx_for(g)
    diag, val, sw, normed = ctx.dmap_kernel(12)
    vd, v, osw, onormed = dmap.sym_kernel(row_ptr, col, dist, 12)
    assert rel_err(normed, onormed) <= 1e-5
    assert rel_err(sw, osw) <= 1e-5
    assert rel_err(diag, vd) <= 1e-5
    assert rel_err(val, v) <= 2e-5
    # the reference's own normalisation check (diffmaps.rs:482-490): sqrt(degrees) is the eigenvector of eigenvalue 1
    S = dmap.kernel_matrix(row_ptr, col, diag, val)
    np.testing.assert_allclose(S @ sw.astype(np.float64), sw, rtol=1e-4)
    ctx.close()


@pytest.mark.parametrize("d", [2, 5])
def test_dmap_layout_matches_oracle_with_the_same_test_matrix(d):
    x, g = strip_graph(3000, 10)
    row_ptr, col, dist = g.get_neighbours()
    omega = np.random.default_rng(11).standard_normal((3000, 20)).astype(np.float32)
    ctx = ctx_for(g, asked_dim=d)
    ctx.dmap_set_test_matrix(omega)
    y = ctx.dmap_init()
    sig = ctx.dmap_singular_values()
    yo, lam, U = dmap.dmap_layout_randomized(row_ptr, col, dist, asked_dim=d, omega=omega)
    np.testing.assert_allclose(sig, lam, atol=2e-4)
    # set_data_box(., 10): centred, largest coordinate 5 (embedder.rs:1376-1408)
    np.testing.assert_allclose(y.mean(axis=0), 0, atol=1e-4)
    assert abs(np.abs(y).max() - 5) < 1e-4
    for c in range(d):
        gap = min(lam[c] - lam[c + 1], lam[c + 1] - lam[c + 2])              # conditioning of singular vector c+1
        corr = np.corrcoef(y[:, c], yo[:, c])[0, 1]
        assert abs(corr) > 1 - 1e-3 / max(gap, 1e-4) * 1e-2, (c, corr, gap)
        s = np.sign(corr)
        assert np.abs(s * y[:, c] - yo[:, c]).max() < 0.05, c                 # box units (coordinates span 10)
    # the layout is installed as the initial embedding: get_embedding returns it, optimize starts from it
    np.testing.assert_array_equal(ctx.get_embedding(), y)
    # deterministic
    np.testing.assert_array_equal(ctx.dmap_init(), y)
    ctx.close()


def test_dmap_seeded_generator_behaves_like_the_reference_range_finder():
    """Without an injected matrix the Gaussian comes from Philox(seed): singular values fall where the CPU
    restatement's do over seeds, and interlace the exact spectrum from below."""
    x, g = strip_graph(3000, 10)
    row_ptr, col, dist = g.get_neighbours()
    vd, v, sw, normed = dmap.sym_kernel(row_ptr, col, dist)
    S = dmap.kernel_matrix(row_ptr, col, vd, v)
    lam = np.sort(np.abs(np.linalg.eigvalsh(S.toarray())))[::-1][:20]
    ref = np.array([dmap.subspace_svd(S, seed=s)[0] for s in range(6)])
    sigs = []
    for seed in (1, 2, 3):
        ctx = ctx_for(g, seed=seed)
        ctx.dmap_init(want_output=False)
        sigs.append(ctx.dmap_singular_values())
        ctx.close()
    sigs = np.array(sigs)
    assert np.all(sigs <= lam + 1e-5)
    assert np.all(np.diff(sigs, axis=1) <= 1e-7)
    lo, hi = ref.min(axis=0), ref.max(axis=0)
    spread = np.maximum(hi - lo, 2e-3)
    assert np.all(sigs.mean(axis=0) > lo - 3 * spread) and np.all(sigs.mean(axis=0) < hi + 3 * spread)
    assert not np.array_equal(sigs[0], sigs[1])


def test_embed_with_dmap_init_runs_end_to_end():
    """Embedder with dmap_init=True and no explicit layout: the device computes the diffusion-map layout
    (embedder.rs:308-345) and optimizes from it (:351-356)."""
    x, _ = workloads.gaussian_mixture(6000, 32, seed=4)
    idx, dist = workloads.knn_exact(x, 10)
    g = A.KGraph.from_knn(idx, dist)
    emb = A.Embedder(g, A.EmbedderParams(dmap_init=True, nb_grad_batch=10, grad_step=1.0, seed=5))
    assert emb.embed() == 1
    y0 = emb.get_initial_embedding()
    assert y0.shape == (6000, 2) and abs(np.abs(y0).max() - 5) < 1e-3
    y = emb.get_embedded()
    assert np.isfinite(y).all() and "dmap_init" in emb.host_timings_ms
    assert all(np.isfinite(c) for c in emb.cross_entropy)


def test_dmap_errors():
    row_ptr, col, dist = random_graph(50, 3, 5, seed=1)
    ctx = A.CudaContext(A.EmbedderParams())
    with pytest.raises(A.AnnembedCudaError) as e:
        ctx.dmap_init()                                    # no graph
    assert e.value.status == 5
    ctx.set_graph_csr(row_ptr, col, dist)
    with pytest.raises(A.AnnembedCudaError) as e:
        ctx.dmap_init()                                    # too small for a rank-20 range finder
    assert e.value.status == 6
    ctx.close()


@pytest.mark.parametrize("d", [2, 3])
def test_dmap_against_golden_fixture(d):
    """The committed vectors (tests/golden/dmap_small.npz, graph of hotpath_small.npz: ragged rows, zero distances,
    duplicated rows): kernel entries and, with the fixture's Gaussian matrix, singular values and layout."""
    import os
    here = os.path.dirname(__file__)
    G = np.load(os.path.join(here, "golden", "hotpath_small.npz"))
    D = np.load(os.path.join(here, "golden", "dmap_small.npz"))
    g = A.KGraph(G["row_ptr"], G["col"], G["dist"])
    ctx = ctx_for(g, asked_dim=d)
    diag, val, sw, normed = ctx.dmap_kernel(12)
    assert rel_err(diag, D["diag"]) <= 1e-5 and rel_err(val, D["val"]) <= 2e-5
    assert rel_err(sw, D["sw"]) <= 1e-5 and rel_err(normed, D["normed"]) <= 1e-5
    ctx.dmap_set_test_matrix(D["omega"])
    y = ctx.dmap_init()
    np.testing.assert_allclose(ctx.dmap_singular_values(), D["sigma"], atol=2e-4)
    ref = D[f"layout_d{d}"]
    for c in range(d):
        s = np.sign(np.dot(y[:, c], ref[:, c]))
        assert np.abs(s * y[:, c] - ref[:, c]).max() < 0.05, c
    ctx.close()
