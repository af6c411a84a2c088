"""Host-side logic: KGraph CSR interchange file, re-indexing, shard ranges, Philox known answers and the
counter-based sampler (host build of the product's __host__ __device__ code, tests/hostsim)."""
import numpy as np
import pytest

import annembed_b200 as A
from annembed_b200.dist import shard_range
from tests.conftest import random_graph
from tests.studies import hostsim_binding as hs
from oracle import oracle


def test_csr_file_roundtrip(tmp_path):
    row_ptr, col, dist = random_graph(50, 1, 7, seed=1)
    ids = np.random.default_rng(0).permutation(50).astype(np.uint64)
    g = A.KGraph(row_ptr, col, dist, ids)
    path = str(tmp_path / "g.csr")
    A.write_csr(path, g)
    g2 = A.read_csr(path)
    np.testing.assert_array_equal(g.row_ptr, g2.row_ptr)
    np.testing.assert_array_equal(g.col, g2.col)
    np.testing.assert_array_equal(g.dist, g2.dist)
    np.testing.assert_array_equal(g.data_id, g2.data_id)
    assert g2.get_max_nbng() == g.get_max_nbng() == 7 and g2.get_nb_nodes() == 50
    assert g2.get_idx_from_dataid(g2.get_data_id_from_idx(17)) == 17
    with open(path, "r+b") as f:
        f.write(b"XXXX")
    with pytest.raises(ValueError):
        A.read_csr(path)
    with pytest.raises(ValueError):
        A.KGraph(row_ptr, col[:-1], dist)


def test_reindexing_follows_data_ids():
    """get_embedded_reindexed: row i -> row DataId(i) (embedder.rs:397-403)."""
    idx = np.array([[1], [2], [0]])
    g = A.KGraph.from_knn(idx, np.ones((3, 1), np.float32), data_id=np.array([2, 0, 1]))
    e = A.Embedder(g, A.EmbedderParams(dmap_init=False))
    e.embedding = np.array([[0, 0], [1, 1], [2, 2]], np.float32)
    np.testing.assert_array_equal(e.get_embedded_reindexed(), [[1, 1], [2, 2], [0, 0]])
    np.testing.assert_array_equal(e.get_embedded_by_dataid(2), [0, 0])
    g.data_id = np.array([0, 1, 7], np.uint64)
    with pytest.raises(IndexError):
        e.get_embedded_reindexed()
    e2 = A.Embedder(g, A.EmbedderParams())
    with pytest.raises(RuntimeError):
        e2.get_embedded_reindexed()


def test_shard_ranges_partition_nodes():
    for n in (1, 7, 70000, 11_000_000):
        for r in (1, 2, 3, 4, 8):
            ranges = [shard_range(n, k, r) for k in range(r)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(r - 1))
            pad = ((n + r - 1) // r + 31) // 32 * 32
            assert all(hi - lo <= pad and (lo % 32 == 0 or lo == n) for lo, hi in ranges)


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors for philox4x32 with 10 rounds."""
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, exp in kat:
        np.testing.assert_array_equal(hs.philox(ctr, key), np.array(exp, np.uint32))


def test_node_uniforms_are_uniform_and_uncorrelated():
    """The systematic-sampling offsets u_i(epoch) (sgd_core.cuh node_uniform: Philox2x32 epoch key + 2-round
    multiply-xorshift over the node id): uniform over nodes and over epochs, no serial correlation either way,
    24-bit grid, deterministic, seed-sensitive."""
    from scipy import stats
    n = 200000
    u = hs.node_uniforms(0, n, epoch=7, seed=3)
    assert u.min() >= 0.0 and u.max() < 1.0
    assert np.all(u * 16777216.0 == np.floor(u * 16777216.0))
    assert stats.kstest(u, "uniform").pvalue > 1e-3
    for lag in (1, 2, 3, 4, 32, 1024):                                        # neighbouring nodes / lanes / tiles
        assert abs(np.corrcoef(u[:-lag], u[lag:])[0, 1]) < 4.0 / np.sqrt(n)
    chi = stats.chisquare(np.bincount((u * 256).astype(np.int64), minlength=256)).pvalue
    assert chi > 1e-3
    # one node (and a group of 4 lanes) over consecutive mini-epochs
    E = 4000
    ue = np.stack([hs.node_uniforms(12344, 4, epoch=e, seed=3) for e in range(E)])   # (E, 4)
    for c in range(4):
        assert stats.kstest(ue[:, c], "uniform").pvalue > 1e-3
        assert abs(np.corrcoef(ue[:-1, c], ue[1:, c])[0, 1]) < 4.0 / np.sqrt(E)
    assert abs(np.corrcoef(ue[:, 0], ue[:, 1])[0, 1]) < 4.0 / np.sqrt(E)         # two nodes across epochs
    # pairs (u_i, u_{i+1}) fill the unit square evenly
    h2 = np.histogram2d(u[0::2], u[1::2], bins=16, range=[[0, 1], [0, 1]])[0].ravel()
    assert stats.chisquare(h2).pvalue > 1e-3
    np.testing.assert_array_equal(u, hs.node_uniforms(0, n, epoch=7, seed=3))
    assert np.mean(u == hs.node_uniforms(0, n, epoch=7, seed=4)) < 1e-3
    assert np.mean(u == hs.node_uniforms(0, n, epoch=8, seed=3)) < 1e-3


def test_sampler_expectation_and_rejection():
    """T6 on the host build: E[count_e] = nbs*E/n/M * p_e ; negatives avoid {i, j} U N(i) (embedder.rs:1246-1252)."""
    row_ptr, col, dist = random_graph(400, 4, 10, seed=5)
    n = 400
    scale, p = oracle.edge_weights(row_ptr, col, dist)
    nbs, M = 10, 5
    kappa = nbs * len(col) / n / M
    tot = np.zeros(len(col))
    src = np.repeat(np.arange(n), np.diff(row_ptr.astype(np.int64)))
    for epoch in range(200):
        c, negs = hs.draws(row_ptr, col, p, nbs, M, seed=9, epoch=epoch)
        tot += c
        assert np.all(np.abs(c.astype(np.float64) - kappa * p) < 1.0 + 1e-4)   # systematic sampling: floor or ceil
        per_node = np.add.reduceat(c.astype(np.int64), row_ptr[:-1].astype(np.int64))
        assert set(np.unique(per_node)) <= {int(np.floor(kappa)), int(np.ceil(kappa))}    # every node: ceil(kappa - u)
        fired = c > 0
        assert np.all(negs[~fired] == 0xFFFFFFFF)
        ng = negs[fired]
        assert ng.max() < n
        assert not np.any(ng == src[fired, None]) and not np.any(ng == col[fired, None])
        for e in np.nonzero(fired)[0][:50]:
            i = src[e]
            assert not np.intersect1d(negs[e], col[int(row_ptr[i]):int(row_ptr[i + 1])]).size
    assert np.abs(tot / 200 - kappa * p).max() < 0.18          # 5 sigma of a Bernoulli mean over 200 epochs
    assert abs((tot / 200).sum() / (kappa * p).sum() - 1) < 0.01
    # determinism and key sensitivity
    c1, n1 = hs.draws(row_ptr, col, p, nbs, M, seed=9, epoch=3)
    c2, n2 = hs.draws(row_ptr, col, p, nbs, M, seed=9, epoch=3)
    c3, n3 = hs.draws(row_ptr, col, p, nbs, M, seed=10, epoch=3)
    np.testing.assert_array_equal(c1, c2); np.testing.assert_array_equal(n1, n2)
    assert np.any(n1 != n3)


def test_negative_uniformity_chi2():
    """Accepted negatives are uniform over the nodes allowed for the sampled edge (embedder.rs:1121,1246-1252)."""
    n = 64
    row_ptr, col, dist = random_graph(n, 3, 5, seed=2)
    scale, p = oracle.edge_weights(row_ptr, col, dist)
    src = np.repeat(np.arange(n), np.diff(row_ptr.astype(np.int64)))
    allowed = np.ones((len(col), n))
    for e in range(len(col)):
        i = src[e]
        allowed[e, i] = 0
        allowed[e, col[int(row_ptr[i]):int(row_ptr[i + 1])]] = 0
    allowed /= allowed.sum(1, keepdims=True)
    hist = np.zeros(n)
    expect = np.zeros(n)
    for epoch in range(300):
        c, negs = hs.draws(row_ptr, col, p, 10, 1, seed=1, epoch=epoch)
        fired = c > 0
        hist += np.bincount(negs[fired].reshape(-1), minlength=n)
        expect += 5 * allowed[fired].sum(0)
    chi2 = ((hist - expect) ** 2 / expect).sum()
    assert chi2 < n + 6 * np.sqrt(2 * n), chi2          # chi2(63): mean 63, sd 11


def test_csv_output_matches_rust_lower_exp_format(tmp_path):
    """N4: `{:.5e}` of Rust (tools/io.rs:35,59): 1.23457e0 / -9.87654e-3, no '+', no padded exponent."""
    from annembed_b200.io import rust_lower_exp, write_csv_array2, write_csv_labeled_array2
    assert rust_lower_exp(1.234567) == "1.23457e0"
    assert rust_lower_exp(-0.00987654) == "-9.87654e-3"
    assert rust_lower_exp(0.0) == "0.00000e0"
    assert rust_lower_exp(123456.7) == "1.23457e5"
    m = np.array([[1.5, -2.25e-4], [3e10, 0.0]], np.float32)
    p = str(tmp_path / "e.csv")
    assert write_csv_array2(p, m) == 1
    assert open(p).read() == "1.50000e0,-2.25000e-4\n3.00000e10,0.00000e0\n"
    write_csv_labeled_array2(p, [7, 9], m)
    assert open(p).read().splitlines()[1] == "9,3.00000e10,0.00000e0"


def _vose_alias(w):
    n = len(w)
    q = np.asarray(w, np.float64) * n / np.sum(w)
    prob = np.ones(n, np.float32); alias = np.arange(n, dtype=np.uint32)
    small = [i for i in range(n) if q[i] < 1.0]; large = [i for i in range(n) if q[i] >= 1.0]
    while small and large:
        s_, l_ = small.pop(), large.pop()
        prob[s_] = q[s_]; alias[s_] = l_
        q[l_] = (q[l_] + q[s_]) - 1.0
        (small if q[l_] < 1.0 else large).append(l_)
    tab = np.empty((n, 2), np.uint32)
    tab[:, 0] = prob.view(np.uint32); tab[:, 1] = alias
    return tab


def test_hubness_sampler_follows_the_weights():
    """K6: negatives drawn through the alias table (shared-sector lookups) follow clamp(in-degree,1,n)/sum
    (embedder.rs:826-833,909-931), up to the rejection of {i, j} U N(i)."""
    n = 256
    row_ptr, col, dist = random_graph(n, 3, 6, seed=8)
    scale, p = oracle.edge_weights(row_ptr, col, dist)
    w = oracle.hubness_weights(row_ptr, col).astype(np.float64)
    w[:8] *= 6.0                                       # a few pronounced hubs
    tab = _vose_alias(w)
    src = np.repeat(np.arange(n), np.diff(row_ptr.astype(np.int64)))
    allowed = np.tile(w, (len(col), 1))
    for e in range(len(col)):
        i = src[e]
        allowed[e, i] = 0
        allowed[e, col[int(row_ptr[i]):int(row_ptr[i + 1])]] = 0
    allowed /= allowed.sum(1, keepdims=True)
    hist = np.zeros(n); expect = np.zeros(n)
    for epoch in range(150):
        c, negs = hs.draws(row_ptr, col, p, 10, 1, seed=3, epoch=epoch, neg_alias=tab)
        fired = c > 0
        hist += np.bincount(negs[fired].reshape(-1), minlength=n)
        expect += 5 * allowed[fired].sum(0)
    assert hist.sum() == expect.sum().round()
    chi2 = ((hist - expect) ** 2 / expect).sum()
    assert chi2 < n + 6 * np.sqrt(2 * n), chi2
    assert np.corrcoef(hist, w)[0, 1] > 0.98


class _RecordingContext:
    """Stands in for CudaContext (the `context=` argument of Embedder): records the call sequence of embed()."""

    def __init__(self, n, d):
        self.calls, self.n, self.d = [], n, d

    def set_graph_csr(self, row_ptr, col, dist):
        self.calls.append("set_graph_csr")

    def edge_weights(self, want_outputs=True):
        self.calls.append("edge_weights")

    def get_hubness_counts(self):
        self.calls.append("get_hubness_counts")
        return np.arange(self.n, dtype=np.uint32)

    def set_neg_weights(self, w):
        self.calls.append("set_neg_weights")
        self.neg_w = np.array(w)

    def set_embedding(self, y):
        self.calls.append("set_embedding")

    def dmap_init(self):
        self.calls.append("dmap_init")
        return np.full((self.n, self.d), 0.5, np.float32)

    def optimize(self, want_ce=True):
        self.calls.append("optimize")
        return (2.0, 1.0)

    def get_embedding(self):
        self.calls.append("get_embedding")
        return np.zeros((self.n, self.d), np.float32)

    def get_stats(self):
        return {"positive_samples": 0}


def test_embed_call_sequence_follows_one_step_embed():
    """one_step_embed (embedder.rs:298-371): dmap_init without a layout -> device diffusion-map layout (:308-345);
    explicit layout -> set_embedding; dmap_init=false -> get_random_init(1.) (:348); hubness weights = clamp(count, 1, n)
    (:826-833).  Host logic only: the device context is a recording stand-in."""
    rp100, c100, d100 = random_graph(100, 3, 5, seed=2)
    ctx = _RecordingContext(100, 2)
    e = A.Embedder(A.KGraph(rp100, c100, d100), A.EmbedderParams(dmap_init=True), context=ctx)
    assert e.embed() == 1
    assert ctx.calls == ["set_graph_csr", "edge_weights", "dmap_init", "optimize", "get_embedding"]
    assert e.get_initial_embedding().shape == (100, 2) and e.cross_entropy == (2.0, 1.0)
    assert "dmap_init" in e.host_timings_ms and "set_embedding" not in e.host_timings_ms

    row_ptr, col, dist = random_graph(60, 3, 5, seed=2)
    g = A.KGraph(row_ptr, col, dist)
    # below 80 nodes the device range finder does not apply: random layout, with a warning
    ctx = _RecordingContext(60, 2)
    e = A.Embedder(g, A.EmbedderParams(dmap_init=True), context=ctx)
    with pytest.warns(RuntimeWarning, match="dmap_init"):
        assert e.embed() == 1
    assert "dmap_init" not in ctx.calls and "set_embedding" in ctx.calls

    ctx = _RecordingContext(60, 2)
    y0 = np.zeros((60, 2), np.float32)
    e = A.Embedder(g, A.EmbedderParams(dmap_init=True, hubness_weighting=True), initial_embedding=y0, context=ctx)
    e.embed()
    assert ctx.calls == ["set_graph_csr", "edge_weights", "get_hubness_counts", "set_neg_weights", "set_embedding",
                         "optimize", "get_embedding"]
    assert ctx.neg_w.min() == 1.0 and ctx.neg_w.max() == 59.0          # clamp(count, 1, n)

    ctx = _RecordingContext(60, 3)
    e = A.Embedder(g, A.EmbedderParams(dmap_init=False, asked_dim=3, seed=7), context=ctx)
    e.embed()
    assert "dmap_init" not in ctx.calls and "set_embedding" in ctx.calls
    y_init = e.get_initial_embedding()
    assert y_init.shape == (60, 3) and np.abs(y_init).max() <= 0.5      # uniform in [-0.5, 0.5]^d


def test_wide_multiply_shift_is_exact_and_unbiased_at_1e8():
    """philox.cuh below40: floor(((w << 8 | b) * n) / 2^40) exactly (Python integers), and the bias the 32-bit map has at
    n = 10^8 (bins of 42 or 43 words: 2.3 %) is gone: over ALL 2^40 inputs every bin of below40 holds
    floor or ceil of 2^40 / n = 10995 or 10996 inputs (1e-4).  below_auto switches at 2^20."""
    rng = np.random.default_rng(11)
    for n in (1, 7, 1 << 20, (1 << 20) + 1, 25_000_000, 100_000_000, (1 << 32) - 1):
        w = rng.integers(0, 1 << 32, size=20000, dtype=np.uint64).astype(np.uint32)
        w[:4] = [0, 1, 0xFFFFFFFF, 0xFFFFFFFE]
        b = rng.integers(0, 1 << 32, size=20000, dtype=np.uint64).astype(np.uint32)
        o32, o40, oa = hs.below(w, b, n)
        ref32 = [(int(x) * n) >> 32 for x in w]
        ref40 = [(((int(x) << 8) | (int(y) & 0xFF)) * n) >> 40 for x, y in zip(w, b)]
        assert o32.tolist() == ref32 and o40.tolist() == ref40
        assert oa.tolist() == (ref40 if n > (1 << 20) else ref32)
        assert int(o40.max()) < n
    # bin sizes: inputs x in [0, 2^40) map to floor(x n / 2^40); bin k holds ceil((k+1) 2^40 / n) - ceil(k 2^40 / n) inputs
    n = 100_000_000
    k = np.arange(0, n, 9973, dtype=object)
    size40 = [(-(-(int(i) + 1) * (1 << 40) // n)) - (-(-int(i) * (1 << 40) // n)) for i in k]
    size32 = [(-(-(int(i) + 1) * (1 << 32) // n)) - (-(-int(i) * (1 << 32) // n)) for i in k]
    assert set(size40) <= {(1 << 40) // n, (1 << 40) // n + 1} and max(size40) / min(size40) - 1 < 1e-4
    assert max(size32) / min(size32) - 1 > 0.02


@pytest.mark.parametrize("n,G", [(20003, 16), (20003, 8), (37, 16), (16, 16), (5, 8)])
def test_grouped_alias_tables_encode_the_node_law(n, G):
    """annembed_b200/csrc/alias_tables.hpp (hubness sampler, embedder.rs:909-931 restated per sector / per line): the node
    law the tables encode, computed EXACTLY from their entries, is w / sum(w).
    Sector table: P(i) = 1/nsec * sum_s [prob_s * cond_s(i) + (1 - prob_s) * cond_alias(s)(i)], cond from the thresholds.
    Line tables : P(i) = 1/nl * sum_l [prob_l * inner_l(i) + (1 - prob_l) * inner_alias(l)(i)], inner_L(r) = 1/G * sum_c
    [thr_c * (c == r) + (1 - thr_c) * (alias_c == r)].  Also: the alias sector's / line's data embedded in an entry are that
    sector's / line's own, padding nodes have probability 0."""
    rng = np.random.default_rng(n + G)
    w = np.clip(rng.zipf(1.7, n), 1, n).astype(np.float32)          # heavy-tailed, like in-degrees
    w[rng.integers(0, n, size=max(1, n // 50))] = 0.0                # a few nodes that must never be drawn
    if n == 37:
        w[16:32] = 0.0                                               # a whole line of weight zero
    law = w.astype(np.float64) / w.astype(np.float64).sum()
    sec, t1, t2 = hs.alias_tables(w, G)
    f = lambda a: a.view(np.float32).astype(np.float64)
    # ---- sector level
    nsec = len(sec)
    prob, alias = f(sec[:, 0].copy()), sec[:, 1].astype(np.int64)
    thr_own = np.stack([f(sec[:, 2].copy()), f(sec[:, 3].copy()), f(sec[:, 4].copy())], 1)
    thr_al = np.stack([f(sec[:, 5].copy()), f(sec[:, 6].copy()), f(sec[:, 7].copy())], 1)
    np.testing.assert_array_equal(thr_al, thr_own[alias])
    cond = np.diff(np.concatenate([np.zeros((nsec, 1)), thr_own, np.ones((nsec, 1))], 1), axis=1)     # [nsec, 4]
    assert (cond >= 0).all()
    P = np.zeros(nsec * 4)
    np.add.at(P, (np.arange(nsec)[:, None] * 4 + np.arange(4)).ravel(), (prob[:, None] * cond).ravel() / nsec)
    np.add.at(P, (alias[:, None] * 4 + np.arange(4)).ravel(), ((1 - prob)[:, None] * cond[alias]).ravel() / nsec)
    assert np.abs(P[n:]).max(initial=0.0) == 0.0
    # resolution: the thresholds are fp32 numbers in [0, 1] (2^-24 of the SECTOR's mass per row: a weight-1 node behind a
    # weight-10^4 hub of its sector is off by up to 10^-3 of its own tiny probability, everything else by 10^-6)
    mass4 = np.pad(law, (0, nsec * 4 - n)).reshape(nsec, 4).sum(1).repeat(4)[:n]
    assert (np.abs(P[:n] - law) <= 2e-6 * law + 2.5e-7 * mass4).all()
    assert np.quantile(np.abs(P[:n] - law) / np.maximum(law, 1e-300), 0.99) < 2e-5
    # ---- line level
    nl = len(t1)
    probl, al = f(t1[:, 0].copy()), t1[:, 1].astype(np.int64)
    np.testing.assert_array_equal(t2[:, G:], t2[al][:, :G])
    thr = (t2[:, :G] >> 4).astype(np.float64) / 2.0 ** 24
    ali = (t2[:, :G] & 15).astype(np.int64)
    assert thr.max() <= 1.0 and ali.max() < G
    inner = np.zeros((nl, G))
    np.add.at(inner, (np.arange(nl)[:, None].repeat(G, 1), np.arange(G)[None, :].repeat(nl, 0)), thr / G)
    np.add.at(inner, (np.arange(nl)[:, None].repeat(G, 1), ali), (1 - thr) / G)
    np.testing.assert_allclose(inner.sum(1), 1.0, rtol=1e-12)
    P = np.zeros(nl * G)
    np.add.at(P, (np.arange(nl)[:, None] * G + np.arange(G)).ravel(), (probl[:, None] * inner).ravel() / nl)
    np.add.at(P, (al[:, None] * G + np.arange(G)).ravel(), ((1 - probl)[:, None] * inner[al]).ravel() / nl)
    assert np.abs(P[n:]).max(initial=0.0) == 0.0
    assert (P[:n][w == 0] == 0).all()
    massG = np.pad(law, (0, nl * G - n)).reshape(nl, G).sum(1).repeat(G)[:n]        # accept thresholds: 2^-24 of qi = G w_i / W_line
    assert (np.abs(P[:n] - law) <= 2e-6 * law + 2.5e-7 * massG).all()
    assert np.quantile(np.abs(P[:n] - law) / np.maximum(law, 1e-300), 0.99) < 2e-5


@pytest.mark.parametrize("n", [1, 7, 20003, 200001])
def test_node_alias_table_encodes_the_node_law(n):
    """Node-level table of annembed_cuda_set_neg_weights (≙ WeightedAliasIndex over clamp(in_degree, 1, N), embedder.rs:916-919):
    P(i) = 1/n * [prob_i + sum_{s: alias(s) = i} (1 - prob_s)] computed from the entries is w / sum(w) to fp32 resolution of
    the accept probabilities; zero-weight nodes are never drawn.  n = 200001 takes the multi-threaded path of the builder."""
    rng = np.random.default_rng(n)
    w = np.clip(rng.zipf(1.7, n), 1, n).astype(np.float32)
    if n > 7:
        w[rng.integers(0, n, size=n // 50)] = 0.0
    law = w.astype(np.float64) / w.astype(np.float64).sum()
    tab = hs.node_alias_table(w)
    prob = tab[:, 0].copy().view(np.float32).astype(np.float64)
    alias = tab[:, 1].astype(np.int64)
    assert prob.min() >= 0.0 and prob.max() <= 1.0 and alias.max() < n
    P = prob / n
    np.add.at(P, alias, (1.0 - prob) / n)
    assert (P[w == 0] == 0).all()
    assert np.abs(P.sum() - 1.0) < 1e-9
    # an accept probability is an fp32 number in [0, 1]: 2^-24 / n absolute per entry that points at a node
    cnt = np.bincount(alias, minlength=n) + 1
    assert (np.abs(P - law) <= 1e-7 * cnt / n + 1e-12).all()


def test_cpp_mirror_reads_and_writes_the_csr_file(tmp_path):
    """N4: the ANNKGCSR file through the C++ mirror (include/annembed_embedder.hpp read_csr / write_csr; CPU-only driver
    tests/cpp/test_kgraph_io.cpp): a file written by annembed_b200/kgraph.py is read, checked (rows ascending,
    kgraph.rs:508-509) and written back byte for byte; truncated files, a wrong magic and row_ptr[n] != E are refused."""
    import os
    import subprocess
    from annembed_b200.kgraph import KGraph, read_csr, write_csr
    here = os.path.join(os.path.dirname(__file__), "cpp")
    subprocess.run(["make", "-C", here, "-s", "test_kgraph_io"], check=True)
    exe = os.path.join(here, "test_kgraph_io")
    row_ptr, col, dist = random_graph(300, 1, 9, seed=5, zero_frac=0.1)
    ids = np.random.default_rng(1).permutation(300).astype(np.uint64) + 1000
    g = KGraph(row_ptr, col, dist, ids, max_nbng=9)
    a, b = str(tmp_path / "a.csr"), str(tmp_path / "b.csr")
    write_csr(a, g)
    out = subprocess.run([exe, a, b], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip() == f"n=300 E={len(col)} max_nbng=9 first_id={int(ids[0])}"
    assert open(a, "rb").read() == open(b, "rb").read()
    g2 = read_csr(b)
    np.testing.assert_array_equal(g2.col, col); np.testing.assert_array_equal(g2.data_id, ids)
    raw = open(a, "rb").read()
    bad = str(tmp_path / "bad.csr")
    broken_row_ptr = bytearray(raw); broken_row_ptr[40 + 8 * 300] ^= 1          # low byte of row_ptr[n]
    for blob in (raw[:-5], b"ANNKGCSX" + raw[8:], raw[:8] + (2).to_bytes(4, "little") + raw[12:], bytes(broken_row_ptr)):
        open(bad, "wb").write(blob)
        out = subprocess.run([exe, bad, "bad"], capture_output=True, text=True)
        assert out.returncode == 0 and out.stdout.startswith("error:"), out.stdout
