"""Synthetic workloads of BASELINE.json `configs`: data, kNN graphs, initial layouts.

Input tooling shared by tests/ and bench.py.  Nothing here is on the timed path: the kNN graph and the initial
layout are built once and fed identically to the CUDA path and to the CPU oracle.
"""
from __future__ import annotations

import numpy as np
import torch


def gaussian_mixture(n: int, dim: int, n_clusters: int = 10, seed: int = 0, sub_dim: int = 20, intrinsic: int = 5,
                     spread: float = 60.0, sigma: float = 2.0, lo: float = 0.0, hi: float = 255.0,
                     anisotropic: bool = False) -> tuple[np.ndarray, np.ndarray]:
    """C1/C2 'MNIST shape' (n x dim, values in [lo,hi]): a mixture of n_clusters Gaussians whose centres lie in a
    random sub_dim-dimensional subspace; each cluster is a low-rank Gaussian (`intrinsic` latent directions with
    decaying scales, like the few degrees of freedom of a handwritten digit) plus small isotropic pixel noise
    `sigma`, so that nearest neighbours carry structure a 2-D layout can preserve.

    anisotropic=True gives the C2 'Fashion shape': more latent directions, unequal cluster sizes and overlapping
    pairs of clusters (denser hubs)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    basis = np.linalg.qr(rng.standard_normal((dim, sub_dim)))[0].astype(np.float32)          # dim x sub_dim
    centres_sub = rng.uniform(-1.0, 1.0, size=(n_clusters, sub_dim)).astype(np.float32) * (hi - lo) * 0.5
    if anisotropic:
        for c in range(1, n_clusters, 2):                                                     # overlapping pairs
            centres_sub[c] = centres_sub[c - 1] + rng.standard_normal(sub_dim).astype(np.float32) * spread * 0.5
        prob = rng.dirichlet(np.full(n_clusters, 3.0))
        labels = rng.choice(n_clusters, size=n, p=prob)
        intrinsic = intrinsic + 3
    else:
        labels = rng.integers(0, n_clusters, size=n)
    centres = centres_sub @ basis.T + (hi + lo) * 0.5
    x = centres[labels].astype(np.float32)
    for c in range(n_clusters):
        sel = np.nonzero(labels == c)[0]
        a = np.linalg.qr(rng.standard_normal((dim, intrinsic)))[0].astype(np.float32)         # dim x intrinsic
        scales = (spread * 0.7 ** np.arange(intrinsic)).astype(np.float32)
        z = rng.standard_normal((len(sel), intrinsic)).astype(np.float32) * scales
        x[sel] += z @ a.T
    x += rng.standard_normal((n, dim)).astype(np.float32) * np.float32(sigma)
    np.clip(x, lo, hi, out=x)
    return x.astype(np.float32), labels.astype(np.int32)


def knn_exact(x, k: int, device: str | None = None, chunk: int = 4096, dtype=torch.float32) -> tuple[np.ndarray, np.ndarray]:
    """Exact L2 kNN (self excluded), rows ascending.  Returns (idx int64 (n,k), dist float32 (n,k)).
    dtype=torch.float64 makes the neighbour lists reproducible across devices (the fp32 GEMM form of the squared
    distance loses ~1 unit in 1e7 to cancellation, enough to reorder near-ties of 784-dimensional points)."""
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    xt = torch.as_tensor(x, dtype=dtype, device=dev)
    n = xt.shape[0]
    idx_out = torch.empty((n, k), dtype=torch.int64, device=dev)
    d_out = torch.empty((n, k), dtype=torch.float32, device=dev)
    xt = xt - xt.mean(0, keepdim=True)                    # distances are translation invariant; smaller cancellation
    sq = (xt * xt).sum(1)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        d2 = sq[s:e, None] + sq[None, :] - 2.0 * (xt[s:e] @ xt.T)
        d2[torch.arange(e - s, device=dev), torch.arange(s, e, device=dev)] = float("inf")
        v, i = torch.topk(d2, k, dim=1, largest=False, sorted=True)
        idx_out[s:e] = i
        d_out[s:e] = v.clamp_min_(0).sqrt_().to(torch.float32)
    return idx_out.cpu().numpy(), d_out.cpu().numpy()


def csr_from_knn(idx: np.ndarray, dist: np.ndarray) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    n, k = idx.shape
    row_ptr = np.arange(0, (n + 1) * k, k, dtype=np.uint64)
    return row_ptr, idx.reshape(-1).astype(np.uint32), np.ascontiguousarray(dist.reshape(-1), np.float32)


def blocked_knn_graph(n: int, dim: int, k: int, seed: int = 0, block: int = 4096, dup_frac: float = 0.005,
                      device: str | None = None, shuffle: bool = True, host_rng: bool = False) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """C3/C4 'Higgs shape': n x dim standardised mixture, kNN restricted to blocks of `block` points of the same
    sub-cluster (cluster-blocked exact kNN), plus dup_frac exact duplicate rows (zero distances,
    kdumap.rs:163-170).  Node ids are shuffled (HNSW insertion order carries no locality) unless shuffle=False.
    host_rng=True draws every random number with the CPU generator (the same points and permutation on any device:
    the fidelity cases, tests/fidelity_cases.py); the default draws on `device` (bench: 11M points in a second).
    Returns CSR (row_ptr u64, col u32, dist f32)."""
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    gdev = torch.device("cpu") if host_rng else dev
    g = torch.Generator(device=gdev)
    g.manual_seed(seed)
    nblk = (n + block - 1) // block
    idx_all = torch.empty((n, k), dtype=torch.int64, device=dev)
    d_all = torch.empty((n, k), dtype=torch.float32, device=dev)
    grp = 64                                             # blocks per batched matmul
    for b0 in range(0, nblk, grp):
        b1 = min(nblk, b0 + grp)
        s, e = b0 * block, min(n, b1 * block)
        m = e - s
        nb = b1 - b0
        pad = nb * block - m
        centre = torch.randn((nb, 1, dim), device=gdev, generator=g).to(dev) * 1.5
        x = torch.randn((nb, block, dim), device=gdev, generator=g).to(dev) + centre
        ndup = int(block * dup_frac)
        if ndup > 0:
            x[:, block - ndup:, :] = x[:, :ndup, :]      # exact duplicates inside each block
        sq = (x * x).sum(2)
        d2 = sq[:, :, None] + sq[:, None, :] - 2.0 * torch.bmm(x, x.transpose(1, 2))
        ar = torch.arange(block, device=dev)
        d2[:, ar, ar] = float("inf")
        if pad:
            d2[-1, :, block - pad:] = float("inf")       # padding rows of the last block are not neighbours
        v, i = torch.topk(d2, k, dim=2, largest=False, sorted=True)
        i = i + (torch.arange(b0, b1, device=dev) * block)[:, None, None]
        idx_all[s:e] = i.reshape(-1, k)[:m]
        d_all[s:e] = v.reshape(-1, k)[:m].clamp_min_(0).sqrt_()
    if ndup > 0:
        # duplicates are at distance exactly 0 in exact arithmetic; the GEMM formulation leaves ~1e-3 noise
        d_all[d_all < 2e-2] = 0.0
        d_all, order = torch.sort(d_all, dim=1, stable=True)
        idx_all = torch.gather(idx_all, 1, order)
    if shuffle:
        perm = torch.randperm(n, device=gdev, generator=g).to(dev)   # new id of old node i is perm[i]
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(n, device=dev)
        idx_all = perm[idx_all][inv]
        d_all = d_all[inv]
    return csr_from_knn(idx_all.cpu().numpy(), d_all.cpu().numpy())


def set_data_box(y: np.ndarray, box_size: float = 10.0) -> np.ndarray:
    """≙ set_data_box (embedder.rs:1376-1408): centre, then scale so that max |coord| = box_size / 2."""
    y = np.array(y, dtype=np.float32, copy=True)
    y -= y.mean(axis=0, keepdims=True, dtype=np.float64).astype(np.float32)
    mx = float(np.abs(y).max())
    y /= np.float32(mx / (box_size / 2.0))
    return y


def pca_init(x: np.ndarray, d: int, box_size: float = 10.0) -> np.ndarray:
    """A spectral stand-in for the diffusion-map initial layout (embedder.rs:308-345): top-d principal
    components, boxed like the reference boxes its dmap layout (`set_data_box(.., 10.)`, embedder.rs:345)."""
    xc = np.asarray(x, np.float64)
    xc = xc - xc.mean(0, keepdims=True)
    w, v = np.linalg.eigh(xc.T @ xc)                       # deterministic (no randomized range finder)
    v = v[:, ::-1][:, :d]
    v = v * np.sign(v[np.abs(v).argmax(0), np.arange(d)])  # fixed sign: largest component positive
    return set_data_box((xc @ v).astype(np.float32), box_size)


def random_init(n: int, d: int, seed: int = 0, size: float = 1.0) -> np.ndarray:
    """≙ get_random_init (embedder.rs:456-470): uniform in [-size/2, size/2]^d."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.uniform(-size / 2, size / 2, size=(n, d)).astype(np.float32)
