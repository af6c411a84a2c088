"""Quality statistics of an embedding -- restatement of embedder.rs:478-753 (A.8 of SURVEY.md).

TEST INFRASTRUCTURE ONLY.  Differences from the reference, stated in SURVEY.md 8(c):
the radius R_i (distance to the nbng-th nearest embedded neighbour) is EXACT here (scipy cKDTree)
whereas the reference gets it from an HNSW built on the embedded points (embedder.rs:527-554,
hnsw_rs not vendored); quantiles are exact instead of CKMS(0.01).
"""
from __future__ import annotations

import numpy as np

from . import oracle


def kth_neighbour_bruteforce(y, nbng, device=None, chunk=2048):
    """index of the nbng-th nearest embedded neighbour (self excluded) by exact blocked search (any dimension)."""
    import torch

    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    yt = torch.as_tensor(np.ascontiguousarray(y, np.float32), device=dev).double()
    sq = (yt * yt).sum(1)
    out = torch.empty(yt.shape[0], dtype=torch.int64, device=dev)
    for s in range(0, yt.shape[0], chunk):
        e = min(yt.shape[0], s + chunk)
        d2 = sq[s:e, None] + sq[None, :] - 2.0 * (yt[s:e] @ yt.T)
        d2[torch.arange(e - s, device=dev), torch.arange(s, e, device=dev)] = -1.0      # self first
        out[s:e] = torch.topk(d2, nbng + 1, dim=1, largest=False, sorted=True)[1][:, nbng]
    return out.cpu().numpy()


def quality_stats(row_ptr, col, y, nbng, kth_index=None):
    """kth_index: optional precomputed index of every node's nbng-th nearest embedded neighbour (else cKDTree)."""

    row_ptr = np.asarray(row_ptr, np.uint64)
    y = np.ascontiguousarray(y, np.float32)
    n = len(row_ptr) - 1
    # embedder.rs:478-522
    t = oracle.transformed_kgraph(row_ptr, col, y).astype(np.float64)
    # embedder.rs:527-554 + kgraph.rs:167-183: per node, length of the longest of its nbng embedded kNN edges
    if kth_index is None:
        from scipy.spatial import cKDTree
        tree = cKDTree(y.astype(np.float64))
        _, ii = tree.query(y.astype(np.float64), k=nbng + 1, workers=-1)
        kth_index = ii[:, nbng]
    # the radius in the same fp32 arithmetic as the transformed edges (the reference compares f32 distances with `<=`,
    # embedder.rs:659: an original neighbour that IS the nbng-th embedded neighbour counts as a match)
    diff = y - y[kth_index]
    acc = np.zeros(n, np.float32)
    for c in range(y.shape[1]):
        acc = (acc + diff[:, c] * diff[:, c]).astype(np.float32)
    radius = np.sqrt(acc).astype(np.float64)
    deg = np.diff(row_ptr.astype(np.int64))
    rad_e = np.repeat(radius, deg)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = t / rad_e
    match_e = (t <= rad_e).astype(np.int64)
    nodes_match = np.add.reduceat(match_e, row_ptr[:-1].astype(np.int64))
    nb_without_match = int((nodes_match == 0).sum())           # :676-678
    mean_nbmatch = float(nodes_match.sum() / max(1, n - nb_without_match))  # :679-680
    qs = (0.05, 0.25, 0.5, 0.75, 0.85, 0.95)
    finite = np.isfinite(ratio)
    return {
        "nb_without_match": nb_without_match,
        "mean_nbmatch": mean_nbmatch,
        "knn_preservation": float(nodes_match.sum() / deg.sum()),  # SURVEY A.8 proposed definition
        "radius_quantiles": [float(np.quantile(radius, q)) for q in qs],
        "ratio_quantiles": [float(np.quantile(ratio[finite], q)) for q in qs],
        "median_ratio": float(np.quantile(ratio[finite], 0.5)),
        "mean_ratio": float(ratio[finite].mean()),                 # :721-726
        "first_dist": t[row_ptr[:-1].astype(np.int64)],           # -> first_dist.csv :729-735
    }
