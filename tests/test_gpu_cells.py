"""The cell-resident epoch kernel (annembed_b200/csrc/cell_epoch.cuh): one launch runs several mini-epochs of a cell of
nodes with the cell's positions in shared memory.

* one sub-step per launch is, by construction, one mini-epoch of the per-mini-epoch kernels: bit-identical layouts
  (same numbering, same draws, same order of application), on graphs whose edges mostly cross cells (the replayed
  in-edges) and on graphs whose cells are closed (the pushed firing counts);
* several sub-steps per launch agree with the host replay of the same semantics (tests/hostsim), to fp32 rounding;
* the cells are what build_relabelling promises: components of at most cell_nodes nodes are not split.
"""
import numpy as np
import pytest

import annembed_b200 as A
from oracle import oracle
from tests.conftest import random_graph
from tests.studies import hostsim_binding as hs

pytestmark = pytest.mark.gpu

LEGACY = 16   # ANNEMBED_FLAG_LEGACY_EPOCH_KERNELS (implies the bulk-synchronous form)
NO_RELABEL = 4
BULK = 32     # ANNEMBED_FLAG_BULK_SYNCHRONOUS: the cell kernel is the bulk-synchronous form's kernel on one rank


def block_graph(nblocks, bsize, kmin, kmax, seed):
    """Disconnected blocks of `bsize` nodes with random neighbours inside the block, node ids shuffled."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rp, cols, dists = [0], [], []
    for b in range(nblocks):
        r, c, d = random_graph(bsize, kmin, kmax, seed=seed * 1000 + b)
        cols.append(c + b * bsize)
        dists.append(d)
        rp.extend((r[1:] + rp[-1]).tolist())
    row_ptr = np.array(rp, np.uint64)
    col = np.concatenate(cols).astype(np.uint32)
    dist = np.concatenate(dists).astype(np.float32)
    n = nblocks * bsize
    perm = rng.permutation(n)                      # new id of old node i
    inv = np.argsort(perm)
    deg = np.diff(row_ptr.astype(np.int64))
    new_rp = np.concatenate([[0], np.cumsum(deg[inv])]).astype(np.uint64)
    new_col = np.empty_like(col); new_dist = np.empty_like(dist)
    for new_i in range(n):
        o = inv[new_i]
        a, b2 = int(row_ptr[o]), int(row_ptr[o + 1])
        w = int(new_rp[new_i])
        new_col[w:w + b2 - a] = perm[col[a:b2]]
        new_dist[w:w + b2 - a] = dist[a:b2]
    return new_rp, new_col, new_dist


def run(row_ptr, col, dist, y0, hub=False, batches=2, **kw):
    kw["flags"] = kw.get("flags", 0) | BULK
    ctx = A.CudaContext(A.EmbedderParams(dmap_init=False, grad_step=1.0, nb_grad_batch=3, hubness_weighting=hub, **kw))
    ctx.set_graph_csr(row_ptr, col, dist)
    ctx.edge_weights(want_outputs=False)
    if hub:
        ctx.set_neg_weights(oracle.hubness_weights(row_ptr, col))
    ctx.set_embedding(y0)
    ctx.optimize_batches(1, batches)
    y, st = ctx.get_embedding(), ctx.get_stats()
    ctx.close()
    return y, st


# the last four cases have kappa = nbs * mean degree / M below 1 (and below 1/2): the event form of phase A
@pytest.mark.parametrize("d,kmax,hub,M,nbs", [(2, 6, False, 3, 2), (2, 10, True, 2, 3), (3, 8, False, 2, 2), (2, 16, False, 4, 10), (4, 6, True, 3, 1),
                                              (2, 6, False, 5, 1), (2, 6, True, 9, 1), (3, 8, False, 12, 1), (2, 16, False, 40, 3)])
def test_one_substep_per_launch_equals_the_per_mini_epoch_kernels(d, kmax, hub, M, nbs):
    row_ptr, col, dist = random_graph(9000, 2, kmax, seed=81)         # an expander: most edges cross the 4096-node cells
    y0 = np.random.default_rng(8).uniform(-1, 1, size=(9000, d)).astype(np.float32)
    kw = dict(asked_dim=d, seed=21, nb_sampling_by_edge=nbs, mini_epochs_per_batch=M)
    ya, sa = run(row_ptr, col, dist, y0, hub, cell_substeps=1, **kw)
    yb, sb = run(row_ptr, col, dist, y0, hub, flags=LEGACY, **kw)
    assert sa["cell_substeps"] == 1 and sb["cell_substeps"] == 0
    assert sa["n_cells"] >= 3 and sa["cross_cell_edges"] > 0.3 * len(col)
    assert sa["positive_samples"] == sb["positive_samples"] > 0
    assert np.abs(ya - y0).max() > 1e-2
    np.testing.assert_array_equal(ya, yb)


@pytest.mark.parametrize("d,hub,M", [(2, False, 3), (2, True, 3), (4, False, 3), (2, False, 24), (2, True, 12)])
def test_closed_cells_one_substep_equals_the_per_mini_epoch_kernels(d, hub, M):
    bs = 3008 if d <= 2 else 1504                                     # whole tiles, below the cell size (4096 / 2048 nodes)
    row_ptr, col, dist = block_graph(5, bs, 3, 6, seed=5)             # 5 components -> 5 closed cells
    n = 5 * bs
    y0 = np.random.default_rng(9).uniform(-1, 1, size=(n, d)).astype(np.float32)
    kw = dict(asked_dim=d, seed=22, nb_sampling_by_edge=3, mini_epochs_per_batch=M)     # M = 12, 24: kappa below 1, 1/2
    ya, sa = run(row_ptr, col, dist, y0, hub, cell_substeps=1, **kw)
    yb, sb = run(row_ptr, col, dist, y0, hub, flags=LEGACY, **kw)
    assert sa["n_cells"] == 5 and sa["cross_cell_edges"] == 0, (sa["n_cells"], sa["cross_cell_edges"])
    assert sa["positive_samples"] == sb["positive_samples"] > 0
    np.testing.assert_array_equal(ya, yb)
    # the default launch length on closed cells is longer than one sub-step, and it is a different (equally valid)
    # realisation: the negatives are read from the layout at the start of the launch
    yc, sc = run(row_ptr, col, dist, y0, hub, **kw)
    assert sc["cell_substeps"] == min(M, 16) and sc["epoch_launches"] == 2 * ((M + 15) // 16)
    assert sc["positive_samples"] == sa["positive_samples"]
    assert np.isfinite(yc).all() and np.abs(yc - ya).max() > 0


@pytest.mark.parametrize("d,S,kmax,nbs,M", [(2, 2, 6, 2, 2), (2, 3, 10, 3, 3), (4, 2, 6, 2, 2), (2, 4, 6, 1, 8), (2, 3, 10, 1, 15)])
def test_several_substeps_match_the_host_replay(d, S, kmax, nbs, M):
    """One launch of S sub-steps against tests/hostsim (identity numbering: ANNEMBED_FLAG_NO_RELABEL, fixed grid of cells).
    The dynamics are chaotic, so the comparison is on quantiles as in test_epoch_kernel_matches_host_replay."""
    n = 10000
    row_ptr, col, dist = random_graph(n, 3, kmax, seed=83)
    y0 = np.random.default_rng(3).uniform(-2, 2, size=(n, d)).astype(np.float32)
    ctx = A.CudaContext(A.EmbedderParams(asked_dim=d, dmap_init=False, grad_step=1.0, nb_grad_batch=4, nb_sampling_by_edge=nbs,
                                         mini_epochs_per_batch=M, seed=99, flags=NO_RELABEL | BULK, cell_substeps=S))
    ctx.set_graph_csr(row_ptr, col, dist)
    scale, p = ctx.edge_weights()
    es = ctx.get_embedded_scales()
    ctx.set_embedding(y0)
    ctx.optimize_batches(1, 1)                       # M mini-epochs in launches of S sub-steps
    y, st = ctx.get_embedding(), ctx.get_stats()
    assert st["epoch_launches"] == (M + S - 1) // S and st["cell_substeps"] == S
    cell_nodes = int(st["cell_nodes"])
    assert cell_nodes == (4096 if d <= 2 else 2048)
    y_host, done = hs.optimize_cells(row_ptr, col, p, es, y0, 1.0, 1.0, nbs, 4, M, 99, None, 1, 1, cell_nodes=cell_nodes, substeps=S)
    assert st["positive_samples"] == done
    err = np.abs(y - y_host).max(axis=1)
    assert np.median(err) < 1e-6 and np.quantile(err, 0.98) < 1e-3, (np.median(err), np.quantile(err, 0.98), err.max())
    # and it is NOT the flat (one global snapshot per mini-epoch) semantics: the cells do see their own moves
    y_flat, _ = hs.optimize(row_ptr, col, p, es, y0, 1.0, 1.0, nbs, 4, M, 99, None, 1, 1)
    assert np.quantile(np.abs(y - y_flat).max(axis=1), 0.9) > 1e-4
    ctx.close()


def test_hub_cell_with_a_large_byte_map():
    """A node with thousands of in-edges: its cell's firing-count map is much larger than cell_nodes * k."""
    n = 12000
    row_ptr, col, dist = random_graph(n, 3, 6, seed=85)
    col = col.copy()
    rng = np.random.default_rng(1)
    for i in rng.choice(np.arange(1, n), size=5000, replace=False):   # 5000 nodes get node 0 as their last neighbour
        a, b = int(row_ptr[i]), int(row_ptr[i + 1])
        if 0 not in col[a:b]:
            col[b - 1] = 0
    y0 = np.random.default_rng(2).uniform(-1, 1, size=(n, 2)).astype(np.float32)
    kw = dict(asked_dim=2, seed=5, nb_sampling_by_edge=4, mini_epochs_per_batch=2)
    ya, sa = run(row_ptr, col, dist, y0, cell_substeps=1, **kw)
    yb, sb = run(row_ptr, col, dist, y0, flags=LEGACY, **kw)
    assert sa["cell_substeps"] == 1
    np.testing.assert_array_equal(ya, yb)
