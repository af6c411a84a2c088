"""T1/T3: pins the CPU oracle -- hand-derived known answers, the reference's own two bisection
tests (tools/dichotomy.rs:75-91), C restatement vs the independent numpy restatement, structural properties."""
import math

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle
from tests.conftest import random_graph

F = np.float32


def tiny_graph():
    # node 0: neighbours 1,2,3 at 1,2,3 ; first distances of 1,2,3 are .5, 1.5, 1.0 ; node 4 all-equal row ; node 5 zero row
    row_ptr = np.array([0, 3, 5, 7, 9, 12, 14], np.uint64)
    col = np.array([1, 2, 3, 0, 2, 0, 1, 0, 1, 0, 1, 2, 0, 1], np.uint32)
    dist = np.array([1, 2, 3, .5, 2, 1.5, 2, 1., 4, 2, 2, 2, 0, 0], np.float32)
    return row_ptr, col, dist


def test_edge_weights_hand_values_beta1():
    row_ptr, col, dist = tiny_graph()
    scale, p = oracle.edge_weights(row_ptr, col, dist, 1.0, 1.0)
    # node 0: mean_rho = (.5 + 1.5 + 1 + 1)/4 = 1 -> scale 1 ; w = exp(-(0,1,2))
    assert scale[0] == pytest.approx(1.0, rel=1e-7)
    w = np.exp(-np.array([0.0, 1.0, 2.0]))
    np.testing.assert_allclose(p[0:3], w / w.sum(), rtol=2e-7)
    assert p[0:3].sum() == pytest.approx(1.0, abs=2e-7)
    # node 4: all neighbours at distance 2 -> uniform 1/3 (kdumap.rs:224-230); scale still mean_rho
    np.testing.assert_allclose(p[9:12], [1 / 3] * 3, rtol=1e-7)
    assert scale[4] == pytest.approx((1.0 + .5 + 1.5 + 2.0) / 4, rel=1e-7)
    # node 5: all distances zero -> uniform (kdumap.rs:163-170)
    np.testing.assert_allclose(p[12:14], [.5, .5], rtol=1e-7)


def test_edge_weights_hand_values_beta2_and_scale_rho():
    row_ptr, col, dist = tiny_graph()
    scale, p = oracle.edge_weights(row_ptr, col, dist, 1.0, 2.0)
    w = np.exp(-np.array([0.0, 1.0, 4.0]))
    np.testing.assert_allclose(p[0:3], w / w.sum(), rtol=3e-7)
    scale, p = oracle.edge_weights(row_ptr, col, dist, 0.5, 1.0)
    assert scale[0] == pytest.approx(0.5, rel=1e-7)
    w = np.exp(-np.array([0.0, 2.0, 4.0]))
    np.testing.assert_allclose(p[0:3], w / w.sum(), rtol=3e-7)


def test_edge_weights_floor_and_zero_scale():
    # floor: a far last neighbour gets PROBA_MIN before normalisation (kdumap.rs:183)
    row_ptr = np.array([0, 2, 4, 6], np.uint64)
    col = np.array([1, 2, 0, 2, 0, 1], np.uint32)
    dist = np.array([1, 100, 1, 50, 1, 60], np.float32)
    scale, p = oracle.edge_weights(row_ptr, col, dist, 1.0, 1.0)
    np.testing.assert_allclose(p[0:2], np.array([1.0, 1e-4]) / 1.0001, rtol=1e-6)
    # scale == 0 with a positive last distance: arguments are NaN / inf, f32::max keeps 1e-4 for every edge -> uniform
    row_ptr = np.array([0, 2, 4, 6], np.uint64)
    col = np.array([1, 2, 0, 2, 0, 1], np.uint32)
    dist = np.array([0, 3, 0, 1, 0, 2], np.float32)
    scale, p = oracle.edge_weights(row_ptr, col, dist, 1.0, 1.0)
    assert scale[0] == 0.0
    np.testing.assert_allclose(p, 0.5, rtol=1e-7)


def test_edge_weights_empty_row_is_an_error():
    row_ptr = np.array([0, 1, 1, 2], np.uint64)
    with pytest.raises(ValueError, match="node 1"):
        oracle.edge_weights(row_ptr, np.array([1, 0], np.uint32), np.array([1, 1], np.float32))


def test_c_vs_numpy_restatement_weights(small_graph):
    row_ptr, col, dist = small_graph
    for beta, rho in ((1.0, 1.0), (2.0, 0.75), (0.5, 1.3)):
        s1, p1 = oracle.edge_weights(row_ptr, col, dist, rho, beta)
        s2, p2 = oracle.np_edge_weights(row_ptr, col, dist, rho, beta)
        np.testing.assert_allclose(s1, s2, rtol=1e-6)
        np.testing.assert_allclose(p1, p2, rtol=2e-6, atol=1e-9)


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10_000), c=st.floats(0.1, 50.0), beta=st.sampled_from([0.5, 1.0, 2.0]))
def test_edge_weight_properties(seed, c, beta):
    """T3: sum to 1, non-increasing along the row, p_0 = 1/sum(w), floor, scale equivariance."""
    row_ptr, col, dist = random_graph(60, 2, 9, seed)
    scale, p = oracle.edge_weights(row_ptr, col, dist, 1.0, beta)
    s2, p2 = oracle.edge_weights(row_ptr, col, (dist * F(c)).astype(F), 1.0, beta)
    for i in range(60):
        lo, hi = int(row_ptr[i]), int(row_ptr[i + 1])
        r = p[lo:hi]
        assert abs(float(r.sum()) - 1.0) < 1e-5
        assert np.all(np.diff(r) <= 1e-7)
        assert r.min() >= 1e-4 / (hi - lo) * 0.999
    np.testing.assert_allclose(p2, p, rtol=2e-4, atol=1e-7)       # scaling distances leaves p unchanged
    np.testing.assert_allclose(s2, scale * F(c), rtol=1e-5)       # and scales `scale`


def test_embedded_scales_hand_values():
    s = np.array([1, 1, 1, 1, 0.1, 20], np.float32)       # mean = 4.01666
    out = oracle.embedded_scales(s)
    mean = s.sum() / 6
    np.testing.assert_allclose(out, 0.2 * np.clip(s / mean, 0.25, 4.0), rtol=1e-6)
    assert out[4] == pytest.approx(0.05) and out[5] == pytest.approx(0.8)


def _one_edge_graph():
    row_ptr = np.array([0, 1, 2, 3, 4], np.uint64)
    col = np.array([1, 0, 3, 2], np.uint32)
    return row_ptr, col


def test_sgd_sample_hand_values_b1():
    row_ptr, col = _one_edge_graph()
    p = np.array([.5, .5, .5, .5], np.float32)
    es = np.full(4, .5, np.float32)
    y = np.array([[0, 0], [1, 0], [.49, 1], [5, 5]], np.float32)
    negs = np.array([[2, 2, 2, 2, 2]], np.uint32)
    # gamma = 1: u = 4, coeff = 2/(.25*5) = 1.6, attraction 1.6*(-.5+.5e-4) = -.79992 -> clipped to -.49
    out = oracle.step_fixed(row_ptr, col, p, es, y, 1.0, 1.0, [0], negs[:, :5] * 0 + 3)
    np.testing.assert_allclose(out[1], [.51, 0], atol=1e-6)
    # one negative at (.49,1): d=1, u=4, coeff 1.6, rep 1/16 -> .1 ; y_i -= (0,1)*.1, five times with the moving y_i
    out = oracle.step_fixed(row_ptr, col, p, es, y, 1.0, 1.0, [0], negs)
    yi = np.array([.49, 0.0])
    for _ in range(5):
        d = (yi[1] - 1.0) ** 2
        u = d / .25
        c = min(1.0 * (2 / (.25 * (1 + u))) / max(u * u, 1 / 16), 2.0)
        yi = yi - (np.array([.49, 1.0]) - yi) * c
    np.testing.assert_allclose(out[0], yi, rtol=1e-5)
    first = 0.0 - (1.0 - 0.0) * 0.1
    assert yi[1] < first                                   # moved further than one repulsion
    # small gradient step: no clipping
    out = oracle.step_fixed(row_ptr, col, p, es, y, 1.0, 0.1, [0], negs * 0 + 3)
    np.testing.assert_allclose(out[1], [1 - .079992, 0], atol=1e-6)


def test_sgd_sample_hand_values_b_half_and_coincident():
    row_ptr, col = _one_edge_graph()
    p = np.array([.5, .5, .5, .5], np.float32)
    es = np.full(4, .5, np.float32)
    y = np.array([[0, 0], [1, 0], [40, 40], [40, 40]], np.float32)
    # b = .5: coeff = 2*.5*(1/3)*.5/.25 = .66667 ; attraction .66667*(-.49995) = -.3333
    out = oracle.step_fixed(row_ptr, col, p, es, y, 0.5, 1.0, [0], np.array([[3] * 5], np.uint32))
    assert out[1, 0] == pytest.approx(1 - .33330, abs=2e-4)
    # coincident points: no attraction (d_ij_scaled == 0, embedder.rs:1223) ; coincident negative keeps the
    # previous gradient (stale-gradient quirk, embedder.rs:1286-1297)
    y2 = np.array([[1, 1], [1, 1], [1, 1], [9, 9]], np.float32)
    out = oracle.step_fixed(row_ptr, col, p, es, y2, 1.0, 1.0, [0], np.array([[2] * 5], np.uint32))
    np.testing.assert_array_equal(out, y2)


def test_step_fixed_c_vs_numpy(small_graph):
    row_ptr, col, dist = small_graph
    n = len(row_ptr) - 1
    scale, p = oracle.edge_weights(row_ptr, col, dist)
    es = oracle.embedded_scales(scale)
    rng = np.random.default_rng(3)
    for d, b in ((2, 1.0), (5, 0.5), (15, 1.0)):
        y = rng.uniform(-.5, .5, size=(n, d)).astype(F)
        edges = rng.integers(0, len(col), size=40).astype(np.uint64)
        negs = rng.integers(0, n, size=(40, 5)).astype(np.uint32)
        out = oracle.step_fixed(row_ptr, col, p, es, y, b, 0.7, edges, negs)
        y2 = y.copy()
        src = np.searchsorted(row_ptr, edges, side="right") - 1
        for s in range(40):
            oracle.np_sgd_sample(y2, int(src[s]), int(col[edges[s]]), float(p[edges[s]]), float(es[src[s]]), b, 0.7,
                                 [int(v) for v in negs[s]])
        np.testing.assert_allclose(out, y2, rtol=2e-5, atol=2e-6)


def test_cross_entropy_hand_value_and_numpy(small_graph):
    row_ptr, col = _one_edge_graph()
    p = np.array([.5, .5, .5, .5], np.float32)
    es = np.full(4, .5, np.float32)
    y = np.array([[0, 0], [1, 0], [0, 0], [1, 0]], np.float32)
    one = -.5 * math.log(.2) - .5 * math.log(.8)
    assert oracle.cross_entropy(row_ptr, col, p, es, y, 1.0) == pytest.approx(4 * one, rel=1e-6)
    # coincident points: w clamped to 1 - eps (embedder.rs:1338-1341)
    yc = np.zeros((4, 2), np.float32)
    eps = float(np.finfo(np.float32).eps)
    exp = 4 * (-.5 * math.log(1 - eps) - .5 * math.log(eps))
    assert oracle.cross_entropy(row_ptr, col, p, es, yc, 1.0) == pytest.approx(exp, rel=1e-6)
    row_ptr, col, dist = small_graph
    scale, pw = oracle.edge_weights(row_ptr, col, dist)
    es = oracle.embedded_scales(scale)
    yy = np.random.default_rng(0).uniform(-1, 1, size=(len(row_ptr) - 1, 3)).astype(F)
    for b in (1.0, 0.5):
        assert oracle.cross_entropy(row_ptr, col, pw, es, yy, b) == pytest.approx(
            oracle.np_cross_entropy(row_ptr, col, pw, es, yy, b), rel=1e-9)


def test_reference_dichotomy_known_answers():
    """The reference's own tests: tools/dichotomy.rs:75-91 (sqrt(2) within 1e-4)."""
    (rc1, r1), (rc2, r2) = oracle.dichotomy_reference_tests()
    assert rc1 == 0 and rc2 == 0
    assert abs(r1 - math.sqrt(2)) < 1e-4 and abs(r2 - math.sqrt(2)) < 1e-4


def test_scale_from_umap_solves_normalisation():
    d = np.array([.3, .5, .9, 1.4, 2.0], np.float32)
    rc, s, w = oracle.scale_from_umap(d, 3.5)               # f(1) = 2.88 < 3.5: root in [0,1], converges
    assert rc == 0 and abs(float(w.sum()) - 3.5) < 2e-5 and w[0] == 1.0
    np.testing.assert_allclose(w, np.exp(-(d - d[0]) / s), rtol=1e-5)
    # f(1) > target: the bracket becomes [1, f32::MAX] and 100 halvings are not enough -> Err (dichotomy.rs:59-61),
    # which get_scale_from_umap unwraps (embedder.rs:776): the reference would panic; the restatement reports 1
    rc, _, _ = oracle.scale_from_umap(d, 2.5)
    assert rc == 1


def test_hogwild_loop_improves_layout_and_counts_samples(small_graph):
    row_ptr, col, dist = small_graph
    n = len(row_ptr) - 1
    scale, p = oracle.edge_weights(row_ptr, col, dist)
    es = oracle.embedded_scales(scale)
    y0 = np.random.default_rng(1).uniform(-.5, .5, size=(n, 2)).astype(F)
    y, done = oracle.optimize(row_ptr, col, p, es, y0, 1.0, 1.0, 10, 5, seed=3, n_threads=2)
    assert done == 5 * 10 * len(col) and np.isfinite(y).all()
    # last batch has grad_step 0 (embedder.rs:875): running only it changes nothing
    y2, _ = oracle.optimize(row_ptr, col, p, es, y, 1.0, 1.0, 10, 5, seed=4, first_batch=5, n_batches=1, n_threads=2)
    np.testing.assert_array_equal(y, y2)
    # hubness weights = clamp(in-degree, 1, n)
    w = oracle.hubness_weights(row_ptr, col)
    np.testing.assert_array_equal(w, np.clip(np.bincount(col, minlength=n), 1, n).astype(F))


@pytest.mark.parametrize("d,b,hub", [(2, 1.0, False), (5, 0.5, True), (15, 1.0, False)])
def test_reference_layout_twin_is_the_same_loop(d, b, hub):
    """oracle_optimize_reference_layout (the CPU arm's timing twin: per-node heap rows behind Arc + RwLock, heap copies
    per access, embedder.rs:939-941,1071-1073,1186-1301) applies the same samples with the same arithmetic as the
    plain-array loop: on one thread the two layouts are bit-identical; on several threads both are Hogwild runs of the
    same optimisation (finite, same sample count, cross entropy within the run-to-run spread)."""
    row_ptr, col, dist = random_graph(1500, 3, 9, seed=21, zero_frac=0.05)
    n = len(row_ptr) - 1
    scale, p = oracle.edge_weights(row_ptr, col, dist)
    es = oracle.embedded_scales(scale)
    y0 = np.random.default_rng(2).uniform(-.5, .5, size=(n, d)).astype(F)
    w = oracle.hubness_weights(row_ptr, col) if hub else None
    ya, da = oracle.optimize(row_ptr, col, p, es, y0, b, 1.0, 10, 4, neg_w=w, seed=5, n_threads=1)
    yb, db = oracle.optimize(row_ptr, col, p, es, y0, b, 1.0, 10, 4, neg_w=w, seed=5, n_threads=1, reference_layout=True)
    assert da == db == 4 * 10 * len(col)
    assert np.abs(ya - y0).max() > 0.1
    np.testing.assert_array_equal(ya, yb)
    ces = []
    for rl in (False, True):
        y, done = oracle.optimize(row_ptr, col, p, es, y0, b, 1.0, 10, 4, neg_w=w, seed=6, n_threads=4, reference_layout=rl)
        assert done == da and np.isfinite(y).all()
        ces.append(oracle.cross_entropy(row_ptr, col, p, es, y, b))
    assert abs(ces[0] - ces[1]) < 0.15 * min(ces)


def test_mini_embed_full_on_the_oracle():
    """The reference's only end-to-end test of the path, embedder.rs:1435-1467 (mini_embed_full), run on the oracle: 500
    points of gen_rand_data_f32 (:1419-1433: v = 2 i U * U per coordinate, dimension 20), L1 kNN graph with knbn = 10
    (exact here, HNSW there), EmbedderParams::default() with asked_dim = 5 -- the reference asserts embed().is_ok(); here:
    every weight row is a probability row, the loop finishes with finite coordinates, and graph neighbours end up closer
    than random pairs (the cross entropy is NOT a monotone witness: from a random layout it rises, here as in the bench).
    (tests/cpp/test_embedder.cpp runs the same case through the CUDA path.)"""
    rng = np.random.default_rng(0)
    n, dim, knbn = 500, 20, 10
    val = 2.0 * np.arange(n, dtype=np.float32) * rng.random(n, dtype=np.float32)
    data = val[:, None] * rng.random((n, dim), dtype=np.float32)
    l1 = np.abs(data[:, None, :] - data[None, :, :]).sum(-1)
    np.fill_diagonal(l1, np.inf)
    idx = np.argsort(l1, axis=1, kind="stable")[:, :knbn]
    dist = np.take_along_axis(l1, idx, 1).astype(F)
    row_ptr = (np.arange(n + 1) * knbn).astype(np.uint64)
    col = idx.astype(np.uint32).ravel()
    scale, p = oracle.edge_weights(row_ptr, col, dist.ravel())            # scale_rho 1, beta 1: embedparams.rs:107-132
    np.testing.assert_allclose(p.reshape(n, knbn).sum(1), 1.0, rtol=1e-5)
    assert (p > 0).all() and (scale > 0).all()
    es = oracle.embedded_scales(scale)
    y0 = rng.uniform(-.5, .5, size=(n, 5)).astype(F)                      # get_random_init(1.), embedder.rs:456-470
    y, done = oracle.optimize(row_ptr, col, p, es, y0, 1.0, 2.0, 10, 20, seed=1, n_threads=2)   # defaults: grad_step 2, 20 batches
    assert done == 20 * 10 * len(col) and np.isfinite(y).all()
    assert np.isfinite(oracle.cross_entropy(row_ptr, col, p, es, y, 1.0))
    src = np.repeat(np.arange(n), knbn)
    d_edges = np.linalg.norm(y[src] - y[col], axis=1)
    d_rand = np.linalg.norm(y[rng.integers(0, n, 5000)] - y[rng.integers(0, n, 5000)], axis=1)
    assert np.median(d_edges) < 0.5 * np.median(d_rand), (np.median(d_edges), np.median(d_rand))


def test_transformed_kgraph_running_minimum():
    row_ptr = np.array([0, 3, 4, 5, 6], np.uint64)
    col = np.array([1, 2, 3, 0, 0, 0], np.uint32)
    y = np.array([[0, 0], [3, 0], [1, 0], [2, 0]], np.float32)
    t = oracle.transformed_kgraph(row_ptr, col, y)
    # distances in graph order 3,1,2 -> running minimum 3,1,1 -> sorted 1,1,3 (embedder.rs:500-512)
    np.testing.assert_allclose(t[:3], [1, 1, 3])
