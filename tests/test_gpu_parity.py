"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs and against the committed golden vectors.  Tolerances are the north star's (BASELINE.json):
edge weights <= 1e-5 relative (fp32), one fixed-sample step <= 1e-4, cross-entropy <= 1e-6 relative."""
import os

import numpy as np
import pytest

import annembed_b200 as A
from oracle import oracle
from tests.conftest import random_graph
from tests.studies import hostsim_binding as hs

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hotpath_small.npz"))


def ctx_for(row_ptr, col, dist, **kw):
    p = A.EmbedderParams(**kw)
    ctx = A.CudaContext(p)
    ctx.set_graph_csr(row_ptr, col, dist)
    return ctx


def rel_err(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30)))


# ------------------------------------------------------------------ K1 (T2, T3)
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_edge_weights_golden(tag):
    rho, beta = G[f"w_{tag}_params"]
    ctx = ctx_for(G["row_ptr"], G["col"], G["dist"], scale_rho=float(rho), beta=float(beta))
    scale, p = ctx.edge_weights()
    assert rel_err(scale, G[f"w_{tag}_scale"]) <= 1e-6
    assert rel_err(p, G[f"w_{tag}_p"]) <= 1e-5
    np.testing.assert_allclose(ctx.get_perplexity(), G[f"w_{tag}_perplexity"], rtol=2e-5)


@pytest.mark.parametrize("seed,kmin,kmax,beta,rho", [(1, 1, 1, 1.0, 1.0), (2, 2, 16, 1.0, 0.75), (3, 6, 6, 2.0, 1.0),
                                                     (4, 3, 31, 0.5, 1.3), (5, 10, 10, 1.0, 1.0)])
def test_edge_weights_vs_oracle_random_ragged(seed, kmin, kmax, beta, rho):
    row_ptr, col, dist = random_graph(5000, kmin, kmax, seed, zero_frac=0.05, dup_rows=50)
    ctx = ctx_for(row_ptr, col, dist, scale_rho=rho, beta=beta)
    scale, p = ctx.edge_weights()
    s_ref, p_ref = oracle.edge_weights(row_ptr, col, dist, rho, beta)
    assert rel_err(scale, s_ref) <= 1e-6
    assert rel_err(p, p_ref) <= 1e-5
    sums = np.add.reduceat(p.astype(np.float64), row_ptr[:-1].astype(np.int64))
    assert np.abs(sums - 1).max() < 1e-5


def test_edge_weights_degenerate_rows():
    # all-equal row, all-zero row, zero scale with positive last distance, floor active (kdumap.rs:161-234)
    row_ptr = np.array([0, 3, 6, 8, 10, 12, 14], np.uint64)
    col = np.array([1, 2, 3, 0, 2, 3, 3, 4, 2, 4, 2, 3, 0, 1], np.uint32)
    dist = np.array([2, 2, 2, 0, 0, 0, 0, 5, 0, 1, 0, 9, 1, 500], np.float32)
    ctx = ctx_for(row_ptr, col, dist)
    scale, p = ctx.edge_weights()
    s_ref, p_ref = oracle.edge_weights(row_ptr, col, dist)
    np.testing.assert_allclose(scale, s_ref, rtol=1e-6)
    np.testing.assert_allclose(p, p_ref, rtol=1e-5)
    np.testing.assert_allclose(p[:6], 1 / 3, rtol=1e-6)


def test_edge_weights_scale_equivariance_on_device():
    row_ptr, col, dist = random_graph(2000, 4, 12, seed=11)
    s1, p1 = ctx_for(row_ptr, col, dist).edge_weights()
    s2, p2 = ctx_for(row_ptr, col, (dist * np.float32(8.0)).astype(np.float32)).edge_weights()   # power of two: exact
    np.testing.assert_array_equal(p1, p2)
    np.testing.assert_array_equal(s2, s1 * np.float32(8.0))


def test_edge_weights_umap_bisection_vs_oracle():
    row_ptr, col, dist = random_graph(300, 5, 9, seed=21)
    dist = (dist * np.float32(3.0)).astype(np.float32)
    ctx = ctx_for(row_ptr, col, dist)
    norm = 2.0
    scale, w, status = ctx.edge_weights_umap(norm)
    n_checked = 0
    for i in range(300):
        lo, hi = int(row_ptr[i]), int(row_ptr[i + 1])
        rc, s_ref, w_ref = oracle.scale_from_umap(dist[lo:hi], norm)
        if rc < 0:
            assert status[i] == 2
            continue
        assert status[i] == rc
        if rc == 0:
            n_checked += 1
            assert abs(float(w[lo:hi].sum()) - norm) < 3e-5
            np.testing.assert_allclose(scale[i], s_ref, rtol=2e-3)
            np.testing.assert_allclose(w[lo:hi], w_ref, rtol=2e-3, atol=1e-6)
    assert n_checked > 50


# ------------------------------------------------------------------ K2
def test_embedded_scales_vs_oracle():
    row_ptr, col, dist = random_graph(70000, 3, 8, seed=31)
    ctx = ctx_for(row_ptr, col, dist)
    scale, _ = ctx.edge_weights()
    es = ctx.get_embedded_scales()
    # fp64 mean on the device; the reference's sequential fp32 sum differs by ~1e-5 relative at this size
    assert rel_err(es, oracle.embedded_scales(scale, f64_sum=True)) <= 1e-6
    assert rel_err(es, oracle.embedded_scales(scale, f64_sum=False)) <= 1e-4
    ctx2 = ctx_for(G["row_ptr"], G["col"], G["dist"])
    ctx2.edge_weights()
    assert rel_err(ctx2.get_embedded_scales(), G["emb_scale"]) <= 1e-6


# ------------------------------------------------------------------ K3 (T4)
@pytest.mark.parametrize("d", [2, 3, 15])
@pytest.mark.parametrize("b", [1.0, 0.5])
@pytest.mark.parametrize("gs", [1.0, 0.05])
def test_step_fixed_golden(d, b, gs):
    ctx = ctx_for(G["row_ptr"], G["col"], G["dist"], asked_dim=d, b=b)
    ctx.set_edge_weights(G["w_a_scale"], G["w_a_p"])
    ctx.set_embedding(G[f"y0_d{d}"])
    ctx.step_fixed(G[f"edges_d{d}"], G[f"negs_d{d}"], gs)
    y = ctx.get_embedding()
    ref = G[f"step_d{d}_b{b}_g{gs}"]
    # 1e-4 relative to the coordinate scale (layout box is O(1)); 64 chained samples
    assert np.abs(y - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())


def test_step_fixed_vs_oracle_single_samples_and_clips():
    row_ptr, col, dist = random_graph(1000, 4, 10, seed=41)
    scale, p = oracle.edge_weights(row_ptr, col, dist)
    es = oracle.embedded_scales(scale)
    rng = np.random.default_rng(5)
    for d, b, gs, box in ((2, 1.0, 2.0, 0.2), (2, 1.0, 0.3, 10.0), (8, 0.7, 1.0, 1.0), (32, 1.0, 1.0, 1.0)):
        y0 = rng.uniform(-box / 2, box / 2, size=(1000, d)).astype(np.float32)
        ctx = ctx_for(row_ptr, col, dist, asked_dim=d, b=b)
        ctx.set_edge_weights(scale, p)
        for trial in range(20):                      # independent single samples: no error accumulation
            e = rng.integers(0, len(col), size=1).astype(np.uint64)
            ng = rng.integers(0, 1000, size=(1, 5)).astype(np.uint32)
            ctx.set_embedding(y0)
            ctx.step_fixed(e, ng, gs)
            y = ctx.get_embedding()
            ref = oracle.step_fixed(row_ptr, col, p, es, y0, b, gs, e, ng)
            moved = np.abs(ref - y0).max()
            assert np.abs(y - ref).max() <= 1e-4 * max(moved, np.abs(ref).max(), 1e-3)
        # the host build of the same device code agrees with the GPU to fp32 rounding
        e = rng.integers(0, len(col), size=50).astype(np.uint64)
        ng = rng.integers(0, 1000, size=(50, 5)).astype(np.uint32)
        ctx.set_embedding(y0)
        ctx.step_fixed(e, ng, gs)
        np.testing.assert_allclose(ctx.get_embedding(), hs.step_fixed(row_ptr, col, p, es, y0, b, gs, e, ng),
                                   rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------ K5 (T5)
@pytest.mark.parametrize("d", [2, 3, 15])
@pytest.mark.parametrize("b", [1.0, 0.5])
def test_cross_entropy_golden(d, b):
    ctx = ctx_for(G["row_ptr"], G["col"], G["dist"], asked_dim=d, b=b)
    ctx.set_edge_weights(G["w_a_scale"], G["w_a_p"])
    ctx.set_embedding(G[f"y0_d{d}"])
    ce = ctx.cross_entropy()
    assert abs(ce - float(G[f"ce_d{d}_b{b}"])) <= 1e-6 * abs(float(G[f"ce_d{d}_b{b}"]))


def test_cross_entropy_large_vs_oracle():
    row_ptr, col, dist = random_graph(50000, 6, 6, seed=51)
    ctx = ctx_for(row_ptr, col, dist)
    scale, p = ctx.edge_weights()
    y = np.random.default_rng(2).uniform(-5, 5, size=(50000, 2)).astype(np.float32)
    ctx.set_embedding(y)
    ref = oracle.cross_entropy(row_ptr, col, p, ctx.get_embedded_scales(), y, 1.0)
    assert abs(ctx.cross_entropy() - ref) <= 1e-6 * abs(ref)


# ------------------------------------------------------------------ sampler (T6) and K4
def test_device_draws_match_host_build_bit_for_bit():
    row_ptr, col, dist = random_graph(3000, 3, 12, seed=61)
    # flags=4 (ANNEMBED_FLAG_NO_RELABEL): the draws are keyed by the internal node ids; the host build knows the caller's
    ctx = ctx_for(row_ptr, col, dist, nb_sampling_by_edge=10, mini_epochs_per_batch=7, seed=1234567890123, flags=4)
    scale, p = ctx.edge_weights()
    for epoch in (0, 1, 77):
        c_dev, n_dev = ctx.debug_draws(epoch)
        c_host, n_host = hs.draws(row_ptr, col, p, 10, 7, 1234567890123, epoch)
        np.testing.assert_array_equal(c_dev, c_host)
        np.testing.assert_array_equal(n_dev, n_host)
    np.testing.assert_array_equal(ctx.get_hubness_counts(), np.bincount(col, minlength=3000))


# flags 4 = no relabelling (debug_draws reports the caller's ids), 64 = ANNEMBED_FLAG_NODE_ALIAS; M = mini_epochs_per_batch
# (0: the default schedule, whose event kernels draw through the line-level tables); G = nodes that share a negative stream
@pytest.mark.parametrize("flags,M,G", [(4, 1, 4), (4 | 64, 1, 4), (4, 0, 16)])
def test_hubness_negatives_follow_the_weights_chi_square(flags, M, G):
    """Hubness sampler (embedder.rs:909-931): the accepted negatives follow clamp(in-degree, 1, n) / sum.  Chi-square of the
    device draws against the law, per lane of a group (the lanes of a group share the sector / line draw, so only the
    draws of ONE lane are independent of each other), for the sector-level alias table (shared sector draw, per-lane row
    from a rotated uniform), the node-level table (flag 64) and the line-level tables of the event kernels (alias method
    over the 16-node lines, alias method inside the line).  (With the internal relabelling the tables are built over the
    relabelled weights; tests/test_gpu_fidelity.py[c3s-True] covers that path end to end.)"""
    from scipy import stats
    n = 20000
    row_ptr, col, dist = random_graph(n, 4, 12, seed=62)
    rng = np.random.default_rng(3)
    # heavy-tailed weights (a few hubs), like the in-degrees of a real kNN graph
    w = np.clip(rng.zipf(1.7, n), 1, n).astype(np.float32)
    ctx = ctx_for(row_ptr, col, dist, hubness_weighting=True, mini_epochs_per_batch=M, flags=flags, seed=99)
    ctx.edge_weights(want_outputs=False)
    ctx.set_neg_weights(w)
    origin = np.repeat(np.arange(n), np.diff(row_ptr).astype(np.int64))
    hist = np.zeros((G, n))          # one histogram per lane of the groups
    for epoch in range(40 if M else 160):
        c, negs = ctx.debug_draws(epoch)             # per edge: firing count and the negatives of its first firing
        for r in range(G):
            sel = np.nonzero((c > 0) & ((origin & (G - 1)) == r))[0]
            hist[r] += np.bincount(negs[sel].reshape(-1), minlength=n)
    ctx.close()
    law = w.astype(np.float64) / w.sum()
    hub = int(np.argmax(w))
    order = np.argsort(law)
    assert hist.sum() > 3e6
    for r in range(G):
        N = hist[r].sum()
        e, c = law[order] * N, hist[r][order]
        edges = np.nonzero(np.diff(np.floor(np.cumsum(e) / 50.0)))[0] + 1       # merge rare nodes: expected counts >= 50
        ce, ee = np.add.reduceat(c, np.r_[0, edges]), np.add.reduceat(e, np.r_[0, edges])
        chi = stats.chisquare(ce, ee * ce.sum() / ee.sum())
        assert chi.pvalue > 1e-3, (r, chi.statistic, len(ce), chi.pvalue)
    Nt = hist.sum()
    assert abs(hist[:, hub].sum() / Nt / law[hub] - 1) < 0.02


@pytest.mark.parametrize("flags,shift", [(4, 4), (4 | 256, 2)])   # 256 = ANNEMBED_FLAG_SECTOR_NEGATIVES (groups of 4 nodes)
def test_uniform_negatives_of_the_event_kernels_chi_square(flags, shift):
    """Uniform sampler of the asynchronous event kernels (embedder.rs:1121 restated per group): the 2^shift nodes whose rows
    share a 128-byte line (32-byte sector) of the layout draw the SAME random line per negative slot and each takes a
    different row of it.  Per lane the accepted negatives must be uniform over the nodes (chi-square); within a group the
    first firings of an epoch share their lines and take distinct rows."""
    from scipy import stats
    n = 20000
    row_ptr, col, dist = random_graph(n, 4, 12, seed=63)
    ctx = ctx_for(row_ptr, col, dist, flags=flags, seed=7)       # default schedule, dimension 2
    ctx.edge_weights(want_outputs=False)
    origin = np.repeat(np.arange(n), np.diff(row_ptr).astype(np.int64))
    G = 1 << shift
    hist = np.zeros((G, n))
    shared = distinct = pairs = 0
    for epoch in range(120):
        c, negs = ctx.debug_draws(epoch)
        fire = np.nonzero(c > 0)[0]
        for r in range(G):
            sel = fire[(origin[fire] & (G - 1)) == r]
            hist[r] += np.bincount(negs[sel].reshape(-1), minlength=n)
        # first firing edge of every firing node (its firing index 0): group-mates share the stream (group, 0, epoch)
        first = fire[np.r_[True, origin[fire][1:] != origin[fire][:-1]]]
        grp = origin[first] >> shift
        same = np.nonzero(grp[1:] == grp[:-1])[0]
        a, b = negs[first[same]], negs[first[same + 1]]
        ok = (a != 0xFFFFFFFF) & (b != 0xFFFFFFFF)
        shared += int(((a >> shift) == (b >> shift))[ok].sum())
        distinct += int((a != b)[ok].sum())
        pairs += int(ok.sum())
    ctx.close()
    assert pairs > 10000
    assert shared >= 0.99 * pairs            # the exceptions are the redraws after a rejection (own row / own node)
    assert distinct >= 0.999 * pairs
    for r in range(G):
        N = hist[r].sum()
        chi = stats.chisquare(hist[r])
        assert N > 20 * n / G and chi.pvalue > 1e-4, (r, N, chi.statistic, chi.pvalue)


BULK = 32    # ANNEMBED_FLAG_BULK_SYNCHRONOUS: the deterministic snapshot kernels (the default on one rank is the asynchronous sweep)


@pytest.mark.parametrize("d,hub,flags", [(2, False, 4 | BULK), (2, False, 4 | 8), (2, True, 4 | BULK), (5, False, 4 | BULK), (15, False, 4 | BULK), (15, False, 4 | 8)])
def test_epoch_kernel_matches_host_replay(d, hub, flags):
    """Bulk-synchronous K4 against the host build of the same mini-epoch body: same draws, same order, fp32 rounding apart.
    ONE mini-epoch is compared: the dynamics are chaotic (repulsion coefficients up to 2 triple a perturbation per
    close negative), so rounding differences between nvcc's fma contraction and the host build grow afterwards."""
    row_ptr, col, dist = random_graph(4000, 3, 10, seed=71)
    n = 4000
    # nb_sampling_by_edge = 1 and one mini-epoch per batch: a batch is exactly one launch of K4
    # flags: 4 = caller's node order kept inside the optimizer (the host build has no relabelling), 8 = in-edge
    # decisions replayed (the multi-rank kernel) instead of pushed by the out-edge kernel
    kw = dict(asked_dim=d, nb_grad_batch=4, nb_sampling_by_edge=1, mini_epochs_per_batch=1, grad_step=1.0, seed=99,
              hubness_weighting=hub, flags=flags)
    ctx = ctx_for(row_ptr, col, dist, **kw)
    scale, p = ctx.edge_weights()
    es = ctx.get_embedded_scales()
    if hub:
        ctx.set_neg_weights(oracle.hubness_weights(row_ptr, col))
    y0 = np.random.default_rng(1).uniform(-2, 2, size=(n, d)).astype(np.float32)
    ctx.set_embedding(y0)
    ctx.optimize_batches(1, 1)                       # 1 mini-epoch at gamma = 0.75
    y = ctx.get_embedding()
    st = ctx.get_stats()
    assert st["epoch_launches"] == 1 and st["mini_epochs_per_batch"] == 1
    assert np.isfinite(y).all() and np.abs(y - y0).max() > 1e-2
    if hub:
        return                                       # host replay needs the device alias table; covered by draws test
    y_host, done = hs.optimize(row_ptr, col, p, es, y0, 1.0, 1.0, 1, 4, 1, 99, None, 1, 1)
    assert st["positive_samples"] == done
    assert abs(done / len(col) - 1) < 0.03           # one batch = nb_sampling_by_edge * E samples in expectation
    err = np.abs(y - y_host).max(axis=1)
    assert np.quantile(err, 0.999) < 1e-4 and np.median(err) < 1e-6, (np.quantile(err, 0.999), np.median(err), err.max())
    # second mini-epoch (batch 2) still agrees on almost every node
    ctx.optimize_batches(2, 1)
    y_host2, _ = hs.optimize(row_ptr, col, p, es, y_host, 1.0, 1.0, 1, 4, 1, 99, None, 2, 1)
    err2 = np.abs(ctx.get_embedding() - y_host2).max(axis=1)
    assert np.quantile(err2, 0.99) < 1e-3, (np.quantile(err2, 0.99), err2.max())


@pytest.mark.parametrize("d,kmax,hub,M", [(2, 6, False, 1), (2, 6, False, 2), (2, 14, True, 1), (3, 8, False, 1), (4, 4, False, 1),
                                          (15, 10, False, 1), (32, 5, False, 1)])
def test_tiled_epoch_kernel_equals_generic_kernel(d, kmax, hub, M):
    """The warp-tiled K4 (k_epoch_out + k_epoch_in) and the thread-per-node K4 run the same per-node program: same
    draws, same firings in the same order (bit-identical after the out-edge phase); the in-edge phase composes the
    same affine maps with a warp scan instead of one after the other, so one mini-epoch agrees to fp32 rounding.
    M = 2 and kmax = 4 give at most 3 firings per node: the path where the 4 lanes of a group exchange their Philox
    blocks by shuffle instead of each computing all of them."""
    row_ptr, col, dist = random_graph(5000, 2, kmax, seed=72)
    y0 = np.random.default_rng(2).uniform(-1, 1, size=(5000, d)).astype(np.float32)
    outs = []
    for flags in (BULK, BULK | 1):                   # 1 = ANNEMBED_FLAG_GENERIC_EPOCH_KERNEL
        # cell_substeps=1: one mini-epoch per launch of the cell kernel, i.e. the generic kernel's global snapshots
        ctx = ctx_for(row_ptr, col, dist, asked_dim=d, nb_grad_batch=3, grad_step=1.0, seed=5, flags=flags,
                      hubness_weighting=hub, nb_sampling_by_edge=1, mini_epochs_per_batch=M, cell_substeps=1)
        ctx.edge_weights(want_outputs=False)
        if hub:
            ctx.set_neg_weights(oracle.hubness_weights(row_ptr, col))
        ctx.set_embedding(y0)
        ctx.optimize_batches(1, 1)                   # M mini-epochs (one batch)
        outs.append((ctx.get_embedding(), ctx.get_stats()["positive_samples"]))
    assert outs[0][1] == outs[1][1]
    assert np.abs(outs[0][0] - y0).max() > 1e-2
    if M == 1:
        np.testing.assert_allclose(outs[0][0], outs[1][0], rtol=1e-5, atol=2e-5)
    else:   # a second mini-epoch amplifies the 1e-7 rounding differences on a few nodes (clipped repulsions)
        err = np.abs(outs[0][0] - outs[1][0]).max(axis=1)
        assert np.quantile(err, 0.999) < 1e-4 and np.median(err) < 1e-6, (np.quantile(err, 0.999), err.max())


@pytest.mark.parametrize("d,kmax,hub,M,nbs", [(2, 6, False, 3, 1), (2, 6, False, 4, 10), (2, 10, True, 2, 3), (4, 8, False, 2, 2), (15, 6, False, 2, 2)])
def test_pushed_firing_counts_equal_replayed_decisions(d, kmax, hub, M, nbs):
    """Single rank: k_epoch_out pushes every fired edge's count into the byte map that k_epoch_in_flags consumes; the
    multi-rank kernel k_epoch_in replays the same decisions from the in-edge records.  Same decisions, same order of
    application: bit-identical layouts over several mini-epochs, and the byte map is left cleared."""
    row_ptr, col, dist = random_graph(6000, 2, kmax, seed=74)
    y0 = np.random.default_rng(4).uniform(-1, 1, size=(6000, d)).astype(np.float32)
    outs = []
    for flags in (BULK, 8):                          # 8 = ANNEMBED_FLAG_REPLAY_IN_EDGES (implies the bulk-synchronous form)
        ctx = ctx_for(row_ptr, col, dist, asked_dim=d, nb_grad_batch=3, grad_step=1.0, seed=6, flags=flags,
                      hubness_weighting=hub, nb_sampling_by_edge=nbs, mini_epochs_per_batch=M)
        ctx.edge_weights(want_outputs=False)
        if hub:
            ctx.set_neg_weights(oracle.hubness_weights(row_ptr, col))
        ctx.set_embedding(y0)
        ctx.optimize_batches(1, 2)                   # 2 batches = 2 M mini-epochs
        outs.append((ctx.get_embedding(), ctx.get_stats()["positive_samples"]))
    assert outs[0][1] == outs[1][1]
    assert np.abs(outs[0][0] - y0).max() > 1e-2
    np.testing.assert_array_equal(outs[0][0], outs[1][0])


def test_relabelling_is_invisible_at_the_boundary():
    """The optimizer renumbers the nodes internally (locality); every array crossing the C ABI stays in the caller's
    order: a batch with gradient step 0 returns the initial layout bit for bit, hubness counts are the caller's
    in-degrees, and the relabelled run is a different random realisation of the same optimisation (close cross entropy)."""
    row_ptr, col, dist = random_graph(5000, 4, 9, seed=75)
    y0 = np.random.default_rng(5).uniform(-1, 1, size=(5000, 2)).astype(np.float32)
    ces = []
    for flags in (0, 4):
        ctx = ctx_for(row_ptr, col, dist, nb_grad_batch=6, grad_step=1.0, seed=11, flags=flags)
        ctx.edge_weights(want_outputs=False)
        np.testing.assert_array_equal(ctx.get_hubness_counts(), np.bincount(col, minlength=5000))
        ctx.set_embedding(y0)
        ctx.optimize_batches(6, 1)                   # last batch: grad_step 0 (embedder.rs:873-876)
        np.testing.assert_array_equal(ctx.get_embedding(), y0)
        ce0, ce1 = ctx.optimize()
        ces.append(ce1)
        assert ce1 < ce0
    assert abs(ces[0] - ces[1]) < 0.05 * ces[1]


def test_rows_longer_than_16_use_the_generic_kernel():
    row_ptr, col, dist = random_graph(3000, 17, 40, seed=73)
    ctx = ctx_for(row_ptr, col, dist, nb_grad_batch=3, grad_step=1.0, mini_epochs_per_batch=60, flags=4 | BULK)
    scale, p = ctx.edge_weights()
    es = ctx.get_embedded_scales()
    y0 = np.random.default_rng(2).uniform(-1, 1, size=(3000, 2)).astype(np.float32)
    ctx.set_embedding(y0)
    ctx.optimize_batches(1, 1)
    st = ctx.get_stats()
    M = st["mini_epochs_per_batch"]
    y_host, done = hs.optimize(row_ptr, col, p, es, y0, 1.0, 1.0, 10, 3, M, 0x5EED, None, 1, 1)
    assert st["positive_samples"] == done and np.isfinite(ctx.get_embedding()).all()


def test_optimize_is_deterministic_and_schedule_matches_reference():
    """The bulk-synchronous form is bit-reproducible for a seed; both forms follow the reference's batch schedule."""
    row_ptr, col, dist = random_graph(3000, 5, 9, seed=81)
    kw = dict(nb_grad_batch=5, grad_step=1.0, seed=7, flags=BULK)
    y0 = np.random.default_rng(3).uniform(-.5, .5, size=(3000, 2)).astype(np.float32)
    outs = []
    for _ in range(2):
        ctx = ctx_for(row_ptr, col, dist, **kw)
        ctx.edge_weights(want_outputs=False)
        ctx.set_embedding(y0)
        ce0, ce1 = ctx.optimize()
        outs.append(ctx.get_embedding())
        assert np.isfinite(outs[-1]).all() and np.isfinite([ce0, ce1]).all()
    np.testing.assert_array_equal(outs[0], outs[1])
    # last batch has grad_step 0 (embedder.rs:873-876): it must not move anything
    before = outs[0]
    ctx.optimize_batches(5, 1)
    np.testing.assert_array_equal(ctx.get_embedding(), before)
    # reset restores the initial layout
    ctx.reset_embedding()
    np.testing.assert_array_equal(ctx.get_embedding(), y0)
    # a different seed gives a different layout
    ctx2 = ctx_for(row_ptr, col, dist, **{**kw, "seed": 8})
    ctx2.edge_weights(want_outputs=False)
    ctx2.set_embedding(y0)
    ctx2.optimize(want_ce=False)
    assert np.abs(ctx2.get_embedding() - before).max() > 1e-3


# ------------------------------------------------------------------ K4, asynchronous form (the default on one rank)
def _async_vs_snapshot(n, d, kmin, kmax, hub, M, flags_async, seed_graph=77):
    row_ptr, col, dist = random_graph(n, kmin, kmax, seed=seed_graph)
    y0 = np.random.default_rng(6).uniform(-1, 1, size=(n, d)).astype(np.float32)
    outs = []
    # flag 4 (no relabelling) on both sides: the draws are keyed by the internal node ids, and the two forms number the
    # nodes differently (random order / locality order)
    # flag 256 (ANNEMBED_FLAG_SECTOR_NEGATIVES): the event kernels share negatives by groups of 4 nodes like the snapshot kernels
    for flags in (flags_async | 4 | 256, BULK | 1 | 4):
        ctx = ctx_for(row_ptr, col, dist, asked_dim=d, nb_grad_batch=2, grad_step=2e-5, seed=13, flags=flags,
                      hubness_weighting=hub, nb_sampling_by_edge=1, mini_epochs_per_batch=M)
        ctx.edge_weights(want_outputs=False)
        if hub:
            ctx.set_neg_weights(oracle.hubness_weights(row_ptr, col))
        ctx.set_embedding(y0)
        ctx.optimize_batches(1, 1)                   # M sweeps / M mini-epochs at gamma = 1e-5
        st = ctx.get_stats()
        outs.append((ctx.get_embedding().astype(np.float64), st["positive_samples"], st["epoch_launches"]))
        ctx.close()
    return y0, outs


# rows: (d, kmin, kmax, hub, M, flags).  kmin == kmax and M == k: kappa == 1 exactly, every node fires once per sweep (the
# pipelined kernel k_sweep_events with full visits); M < k: several firings per visit (k_sweep_async); flags 1: thread per node; flags 128: k_sweep_events_cp
@pytest.mark.parametrize("d,kmin,kmax,hub,M,flags", [(2, 6, 6, False, 6, 0), (3, 8, 8, True, 8, 0), (4, 16, 16, False, 16, 0), (15, 6, 6, False, 6, 0),
                                                     (15, 10, 10, True, 10, 0), (2, 6, 6, False, 6, 128), (3, 8, 8, True, 8, 128),
                                                     (2, 6, 6, False, 2, 0), (2, 3, 10, True, 3, 0), (15, 4, 10, True, 3, 0),
                                                     (2, 6, 6, False, 6, 1), (2, 17, 30, False, 30, 0), (5, 2, 7, True, 2, 1)])
def test_async_sweeps_apply_the_same_samples_as_the_snapshot_kernels(d, kmin, kmax, hub, M, flags):
    """The asynchronous sweep (async_sweep.cuh) draws exactly the samples of the bulk-synchronous mini-epoch with the same
    seed (same firing decisions, same negatives); only the positions a sample reads differ (current instead of the
    snapshot).  With a tiny gradient step every move is small against the distances, so the two agree to first order:
    the difference is a small fraction of the move.  Checks the pipelined kernel, the multi-firing kernel, the wide-row
    paths and the thread-per-node kernel against k_epoch_generic (itself checked against the host build), and that the
    atomic publication loses nothing (sample counts equal, moves equal)."""
    y0, ((ya, sa, la), (yb, sb, lb)) = _async_vs_snapshot(6000, d, kmin, kmax, hub, M, flags)
    assert sa == sb and la == lb == M
    move = np.abs(yb - y0).max(axis=1)
    err = np.abs(ya - yb).max(axis=1)
    assert np.median(move) > 1e-6
    assert np.median(err) < 0.02 * np.median(move) and np.quantile(err, 0.999) < 0.05 * np.quantile(move, 0.999), \
        (np.median(err), np.median(move), np.quantile(err, 0.999), np.quantile(move, 0.999))


@pytest.mark.parametrize("d,kmin,kmax,hub,M", [(2, 6, 6, False, 24), (2, 3, 10, True, 30), (3, 2, 8, False, 9), (15, 6, 6, False, 13), (2, 16, 16, False, 100)])
def test_thinned_sweeps_draw_the_same_law_as_the_snapshot_kernels(d, kmin, kmax, hub, M):
    """Firing probability kappa below 1 per sweep: a tile of 32 nodes fires as a whole with probability kappa (k_sweep_events),
    where the bulk-synchronous mini-epoch lets every node fire with probability kappa.  Different realisations of the
    same law: the sample count is binomial around the same expectation and the moves have the same distribution."""
    n = 20000
    y0, ((ya, sa, la), (yb, sb, lb)) = _async_vs_snapshot(n, d, kmin, kmax, hub, M, 0, seed_graph=79)
    assert la == lb == M
    tiles, kappa = n / 32, sb / (M * n)                       # bulk count: every node, probability kappa, M times
    sigma = 32 * np.sqrt(M * tiles * kappa * (1 - kappa))       # async count: 32 x Binomial(M tiles, kappa)
    assert abs(sa - M * n * kappa) < 5 * sigma + 0.01 * sb, (sa, sb, sigma)
    ma, mb = np.abs(ya - y0).max(axis=1), np.abs(yb - y0).max(axis=1)
    assert np.isfinite(ya).all() and np.median(mb) > 1e-7
    for q in (0.5, 0.9, 0.99):
        assert abs(np.quantile(ma, q) / np.quantile(mb, q) - 1) < 0.1, (q, np.quantile(ma, q), np.quantile(mb, q))


def test_async_runs_are_statistically_reproducible():
    """Two asynchronous runs differ in the interleaving of the warps (like two runs of the reference); they are two
    realisations of the same optimisation: same sample count, cross entropies within 1 %."""
    row_ptr, col, dist = random_graph(20000, 6, 6, seed=78)
    y0 = np.random.default_rng(7).uniform(-1, 1, size=(20000, 2)).astype(np.float32)
    res = []
    for _ in range(2):
        ctx = ctx_for(row_ptr, col, dist, nb_grad_batch=10, grad_step=1.0, seed=3)
        ctx.edge_weights(want_outputs=False)
        ctx.set_embedding(y0)
        ce0, ce1 = ctx.optimize()
        st = ctx.get_stats()
        res.append((ce1, st["positive_samples"], st["mini_epochs_per_batch"], st["epoch_launches"]))
        assert ce1 < ce0 and np.isfinite(ctx.get_embedding()).all()
        ctx.close()
    # default schedule: 4 * nb_sampling_by_edge * mean degree thinned sub-sweeps per batch (firing probability 1/4), 16 per launch
    assert res[0][1] == res[1][1] and res[0][2] == 240 and res[0][3] == 150
    assert abs(res[0][1] / (10 * 10 * len(col)) - 1) < 0.02      # 10 batches x 10 samples per edge (binomial over the firing tiles)
    assert abs(res[0][0] - res[1][0]) < 0.01 * res[0][0]


# ------------------------------------------------------------------ ABI behaviour (T9)
def test_abi_error_codes_and_state_machine():
    row_ptr, col, dist = random_graph(100, 2, 5, seed=91)
    p = A.EmbedderParams()
    ctx = A.CudaContext(p)

    def status(fn, *a):
        with pytest.raises(A.AnnembedCudaError) as e:
            fn(*a)
        return e.value.status

    assert status(ctx.edge_weights) == 5                                   # graph not set
    assert status(ctx.optimize) == 5
    bad = row_ptr.copy(); bad[4] = bad[3]                                   # node 3 empty -> kdumap.rs:75-85
    bad_col = np.delete(col, np.s_[int(row_ptr[3]):int(row_ptr[4])])
    bad_dist = np.delete(dist, np.s_[int(row_ptr[3]):int(row_ptr[4])])
    bad[4:] = row_ptr[4:] - (row_ptr[4] - row_ptr[3])
    assert status(ctx.set_graph_csr, bad, bad_col, bad_dist) == 3
    uns = dist.copy(); lo = int(row_ptr[10]); uns[lo], uns[lo + 1] = uns[lo + 1] + 1, uns[lo]
    assert status(ctx.set_graph_csr, row_ptr, col, uns) == 4
    c2 = col.copy(); c2[0] = 0                                              # self edge of node 0
    assert status(ctx.set_graph_csr, row_ptr, c2, dist) == 1
    c3 = col.copy(); c3[5] = 100
    assert status(ctx.set_graph_csr, row_ptr, c3, dist) == 1
    tiny = (np.array([0, 1, 2, 3], np.uint64), np.array([1, 2, 0], np.uint32), np.ones(3, np.float32))
    assert status(ctx.set_graph_csr, *tiny) == 8                            # nobody acceptable as negative
    ctx.set_graph_csr(row_ptr, col, dist)
    assert status(ctx.optimize) == 5                                        # embedding not set
    ctx.set_embedding(np.zeros((100, 2), np.float32))
    assert status(ctx.optimize) == 5                                        # weights not computed (embedder.rs:802-808)
    ctx.edge_weights(want_outputs=False)
    ctx.optimize()
    # context reuse with another graph
    r2, c2, d2 = random_graph(64, 3, 3, seed=92)
    ctx.set_graph_csr(r2, c2, d2)
    assert status(ctx.get_embedding) == 5
    ctx.edge_weights(want_outputs=False)
    ctx.set_embedding(np.random.default_rng(0).uniform(-1, 1, (64, 2)).astype(np.float32))
    ctx.optimize()
    assert np.isfinite(ctx.get_embedding()).all()
    ctx.close()
    hp = A.EmbedderParams(hubness_weighting=True)
    ctxh = A.CudaContext(hp)
    ctxh.set_graph_csr(row_ptr, col, dist)
    ctxh.edge_weights(want_outputs=False)
    ctxh.set_embedding(np.zeros((100, 2), np.float32))
    assert status(ctxh.optimize) == 5                                       # hubness sampler not provided


def test_embedder_mirror_end_to_end():
    """Embedder::new / embed / get_embedded_reindexed through the host mirror."""
    row_ptr, col, dist = random_graph(2000, 6, 6, seed=93)
    ids = np.random.default_rng(4).permutation(2000).astype(np.uint64)
    g = A.KGraph(row_ptr, col, dist, ids)
    params = A.EmbedderParams(dmap_init=False, nb_grad_batch=6, grad_step=1.0, hubness_weighting=True)
    emb = A.Embedder(g, params)
    assert emb.embed() == 1
    y = emb.get_embedded()
    r = emb.get_embedded_reindexed()
    assert y.shape == (2000, 2) and np.isfinite(y).all()
    np.testing.assert_array_equal(r[ids.astype(np.int64)], y)
    assert emb.cross_entropy[1] < emb.cross_entropy[0]
    assert emb.get_hubness().sum() == len(col)
    # the default dmap_init=True computes the diffusion-map layout on the device (embedder.rs:308-345, tests/test_gpu_dmap.py)
    e2 = A.Embedder(g, A.EmbedderParams(dmap_init=True, nb_grad_batch=3))
    assert e2.embed() == 1 and abs(np.abs(e2.get_initial_embedding()).max() - 5) < 1e-3
    with pytest.raises(A.EmbedError):                       # asked_dim beyond the rank-20 range finder
        A.Embedder(g, A.EmbedderParams(dmap_init=True, asked_dim=25)).embed()


def test_cpp_driver_of_the_host_mirror():
    """T9: the C++ host mirror (include/annembed_embedder.hpp) driven from a C++ program: the reference's
    mini_embed_full (embedder.rs:1435-1467) restated + error behaviour."""
    import subprocess
    here = os.path.join(os.path.dirname(__file__), "cpp")
    subprocess.run(["make", "-C", here, "-s"], check=True)
    out = subprocess.run([os.path.join(here, "test_embedder"), "500"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "test_embedder ok" in out.stdout


def test_in_edge_scan_across_many_rounds_for_a_hub():
    """A node with thousands of in-edges: its affine maps span ~100 sweep rounds of k_epoch_in; the composite carried
    across rounds must equal the one-after-the-other application of the thread-per-node kernel."""
    n, k = 3000, 5
    rng = np.random.default_rng(5)
    row_ptr = np.arange(0, (n + 1) * k, k, dtype=np.uint64)
    col = np.empty((n, k), np.uint32)
    for i in range(n):
        others = rng.choice(n - 2, size=k - 1, replace=False) + 1          # 1..n-2
        others = np.where(others >= i, others + 1, others)
        others = others[others != 7][:k - 1]
        while len(others) < k - 1:
            c = int(rng.integers(1, n))
            if c != i and c != 7 and c not in others:
                others = np.append(others, c)
        col[i] = np.concatenate([[7 if i != 7 else 8], others])             # everybody's first neighbour is node 7
    dist = np.sort(rng.gamma(2.0, 1.0, size=(n, k)).astype(np.float32), axis=1)
    y0 = rng.uniform(-1, 1, size=(n, 2)).astype(np.float32)
    outs = []
    for flags in (BULK, BULK | 1):
        ctx = ctx_for(row_ptr, col.reshape(-1), dist.reshape(-1), nb_grad_batch=3, grad_step=1.0, seed=9, flags=flags,
                      nb_sampling_by_edge=1, mini_epochs_per_batch=1)
        ctx.edge_weights(want_outputs=False)
        assert ctx.get_hubness_counts()[7] == n - 1
        ctx.set_embedding(y0)
        ctx.optimize_batches(1, 1)
        outs.append(ctx.get_embedding())
    assert np.abs(outs[0][7] - y0[7]).max() > 1e-3                          # the hub moved
    np.testing.assert_allclose(outs[0], outs[1], rtol=1e-4, atol=1e-4)      # ~3000 composed maps: fp32 rounding only
    np.testing.assert_allclose(np.delete(outs[0], 7, axis=0), np.delete(outs[1], 7, axis=0), rtol=1e-5, atol=2e-5)
