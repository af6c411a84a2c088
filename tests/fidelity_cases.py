"""The BASELINE.json configs as seeded test cases (graph + initial layout + EmbedderParams) shared by the fidelity
tests (tests/test_gpu_fidelity.py), the fixture generator (tests/golden/make_fidelity_golden.py) and the design
studies.  Everything is seeded; the kNN is exact, so the CPU (this container) and GPU (test box) builds of a case agree
up to distance ties.

c1  : 70k x 784 MNIST-digits shape, k=10, d=2, params examples/mnist_digits.rs:92-100 (30 batches), nbng 50 (:150)
c2  : 70k x 784 Fashion shape (anisotropic mixture), k=10, d=2, params examples/mnist_fashion.rs:92-99 (25 batches)
c3s : C3 shape at reduced size: n x 28 Higgs-shape blocked graph, k=6, 0.5 % duplicates, RANDOM init, scale_rho 0.75,
      40 batches (examples/higgs.rs:204-211,234), nbng 100 (:270)
c4s : the same graph embedded in dimension 15
"""
from __future__ import annotations

import numpy as np
import torch

import workloads

STAT_KEYS = ("nb_without_match", "mean_nbmatch", "knn_preservation", "median_ratio", "mean_ratio")


def summary(q: dict) -> dict:
    return {k: float(q[k]) for k in STAT_KEYS}


def make_case(name: str, n: int | None = None, device: str | None = None) -> dict:
    if name in ("c1", "c2"):
        n = n or 70000
        x, _ = workloads.gaussian_mixture(n, 784, seed=0, anisotropic=(name == "c2"))
        idx, dist = workloads.knn_exact(x, 10, device=device, dtype=torch.float64, chunk=2048)   # fp64: the same lists on any device
        row_ptr, col, dist = workloads.csr_from_knn(idx, dist)
        y0 = workloads.pca_init(x, 2)
        params = dict(asked_dim=2, nb_grad_batch=30 if name == "c1" else 25, scale_rho=1.0, grad_step=1.0)
        return dict(name=name, n=n, k=10, row_ptr=row_ptr, col=col, dist=dist, y0=y0, params=params, nbng=50)
    if name in ("c3s", "c4s"):
        n = n or 1_000_000
        d = 2 if name == "c3s" else 15
        row_ptr, col, dist = workloads.blocked_knn_graph(n, 28, 6, seed=0, device=device, host_rng=True)
        y0 = workloads.random_init(n, d, seed=0)
        params = dict(asked_dim=d, nb_grad_batch=40, scale_rho=0.75, grad_step=1.0)
        return dict(name=name, n=n, k=6, row_ptr=row_ptr, col=col, dist=dist, y0=y0, params=params, nbng=100)
    raise ValueError(name)


def quality_stats(row_ptr, col, y, nbng):
    """oracle/quality.py statistics; the kNN radius by kd-tree in 2-D, by exact blocked search above (a kd-tree in
    dimension 15 degenerates)."""
    from oracle import quality
    y = np.asarray(y)
    kth = quality.kth_neighbour_bruteforce(y, nbng) if y.shape[1] > 3 else None
    return quality.quality_stats(row_ptr, col, y, nbng, kth_index=kth)
