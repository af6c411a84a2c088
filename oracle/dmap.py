"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's diffusion-map initial layout.

PARITY UNPINNED: the reference cannot be built here (no cargo) and holds no numeric test for this path; the random
range finder (`tools/svdapprox.rs:343-425`, unseeded Gaussian test matrix) is replaced by an exact symmetric
eigen-decomposition, which is what the randomized method approximates.  Everything before the SVD is deterministic and
restated operation by operation.

Follows (all under /root/reference/src):
  embedder.rs:308-345      dmap_init branch: DiffusionParams::new(2, Some(5.), Some(12)), alfa 0.5, beta -0.1, then
                           set_data_box(., 10)   (the initial layout has asked_dim columns: intended semantics)
  diffmaps.rs:397-422      laplacian_from_kgraph: nbng = min(gnbn, max_nbng)
  diffmaps.rs:752-849      compute_dmap_nodeparams: L2 scales, mean, zero scales -> mean, two kernel passes (beta < 0)
  diffmaps.rs:1020-1043    get_dist_l2_from_node
  diffmaps.rs:590-679      build_node_param: self edge, all-equal rows, kernel exp(-(d / (sqrt(epsil) sqrt(s_i s_j)))^2), PROBA_MIN floor
  diffmaps.rs:852-942      kernel0_to_density (sparse branch): density proxy q, new scales q^beta * mean_scale
  diffmaps.rs:427-587      compute_laplacian (sparse branch): symmetrisation by max with BOTH (i,j) and (j,i) pushed for every
                           directed entry, alfa density normalisation, D^-1/2 K D^-1/2
  graphlaplace.rs:97-134   do_approx_svd: rank 20, 5 subspace iterations
  diffmaps.rs:1145-1243    embed_from_laplacian: columns 1..asked_dim, (lambda_j / lambda_0)^t U_ij / weight_i, clipped to +-10
  embedder.rs:1376-1408    set_data_box
The reference switches to a dense (P + P^T)/2 formulation below 5000 nodes (graphlaplace.rs:13, diffmaps.rs:445); the
device path and this restatement always use the sparse formulation, which is the one the 11M-node workload runs.
"""
import numpy as np
import scipy.sparse as sp

PROBA_MIN = np.float32(1.0e-4)          # embedder.rs:50


def local_scales(row_ptr, dist, gnbn=12):
    """diffmaps.rs:1020-1043 per node, then :785-797: (scales with zeros replaced by the mean, mean)."""
    row_ptr = np.asarray(row_ptr, np.int64)
    n = len(row_ptr) - 1
    deg = np.diff(row_ptr)
    nbgh = min(gnbn, int(deg.max()))
    d2 = np.asarray(dist, np.float64) ** 2
    s = np.zeros(n)
    for m in range(nbgh):                                   # first nbgh entries of each row
        has = deg > m
        s[has] += d2[row_ptr[:-1][has] + m]
    s = np.sqrt(s / np.maximum(deg, 1))                     # divided by the full row length (:1039)
    mean = s.sum() / n
    s = np.where(s <= 0, mean, s)
    return s.astype(np.float32), np.float32(mean)


def kernel_weights(row_ptr, col, dist, scales, epsil=2.0):
    """diffmaps.rs:590-679 with remap_weight :815-818: (w_self[n], w[E])."""
    row_ptr = np.asarray(row_ptr, np.int64)
    n = len(row_ptr) - 1
    col = np.asarray(col, np.int64)
    dist = np.asarray(dist, np.float32)
    deg = np.diff(row_ptr)
    src = np.repeat(np.arange(n), deg)
    sc = np.asarray(scales, np.float64)
    ls = np.sqrt(sc[col] * sc[src])
    arg = (dist.astype(np.float64) / (np.sqrt(epsil) * ls)) ** 2
    w = np.maximum(np.exp(-arg).astype(np.float32), PROBA_MIN)
    w_self = np.ones(n, np.float32)
    # all-equal rows (:614-627): last strictly positive distance <= first distance, or no positive distance
    last_pos = np.full(n, -1.0)
    pos = dist > 0
    idx = np.nonzero(pos)[0]
    np.maximum.at(last_pos, src[idx], 0)                    # mark rows having a positive distance
    # distances are ascending, so the last positive one is the row maximum
    row_max = np.maximum.reduceat(dist, row_ptr[:-1])
    first = dist[row_ptr[:-1]]
    all_equal = (last_pos < 0) | (row_max <= first)
    p = (1.0 / (deg + 1)).astype(np.float32)
    w_self[all_equal] = p[all_equal]
    ae_edge = all_equal[src]
    w = np.where(ae_edge, p[src], w).astype(np.float32)
    return w_self, w


def symmetrise(row_ptr, col, w):
    """diffmaps.rs:522-539: sym_e = max(w_e, w_reverse(e)) when the reverse entry exists."""
    row_ptr = np.asarray(row_ptr, np.int64)
    n = len(row_ptr) - 1
    src = np.repeat(np.arange(n), np.diff(row_ptr))
    col = np.asarray(col, np.int64)
    key = src * n + col
    rkey = col * n + src
    order = np.argsort(key)
    pos = np.searchsorted(key[order], rkey)
    pos = np.minimum(pos, len(key) - 1)
    found = key[order][pos] == rkey
    w_rev = np.where(found, w[order][pos], 0).astype(np.float32)
    return np.maximum(w, w_rev)


def sym_rowsum(row_ptr, col, sym, diag):
    """row sums of the triplet matrix: every directed entry is pushed as (i,j) and (j,i), the self edge twice."""
    row_ptr = np.asarray(row_ptr, np.int64)
    n = len(row_ptr) - 1
    src = np.repeat(np.arange(n), np.diff(row_ptr))
    q = np.zeros(n)
    np.add.at(q, src, sym.astype(np.float64))
    np.add.at(q, np.asarray(col, np.int64), sym.astype(np.float64))
    return (q + 2.0 * diag).astype(np.float32)


def sym_kernel(row_ptr, col, dist, gnbn=12, alfa=0.5, beta=-0.1, epsil=2.0):
    """The normalised symmetric kernel D^-1/2 K_alfa D^-1/2 as (diag[n], v[E]) over the directed pattern (the matrix is
    diag + A + A^T with A = v on the directed edges), the normaliser sqrt(degrees) and the normed first-pass scales."""
    row_ptr = np.asarray(row_ptr, np.int64)
    n = len(row_ptr) - 1
    deg = np.diff(row_ptr)
    max_nbng = int(deg.max())
    src = np.repeat(np.arange(n), deg)
    col = np.asarray(col, np.int64)
    s1, mean = local_scales(row_ptr, dist, gnbn)
    normed = (s1 / mean).astype(np.float32)
    if beta < 0:
        ws, w = kernel_weights(row_ptr, col, dist, s1, epsil)
        sym = symmetrise(row_ptr, col, w)
        q = sym_rowsum(row_ptr, col, sym, ws).astype(np.float64) / max_nbng          # :928-931
        q = q / (q.sum() / n)
        scales2 = (q ** beta * mean).astype(np.float32)
    else:
        scales2 = np.full(n, mean, np.float32)
    ws, w = kernel_weights(row_ptr, col, dist, scales2, epsil)
    sym = symmetrise(row_ptr, col, w)
    q = sym_rowsum(row_ptr, col, sym, ws).astype(np.float64)
    q = q / (q.sum() / max_nbng)                                                     # :543-545
    v = sym / (q[src] * q[col]) ** alfa
    vd = 2.0 * ws / (q * q) ** alfa
    degrees = np.zeros(n)
    np.add.at(degrees, src, v)
    np.add.at(degrees, col, v)
    degrees += vd
    sw = np.sqrt(degrees)
    v = v / (sw[src] * sw[col])
    vd = vd / (sw * sw)
    return vd.astype(np.float32), v.astype(np.float32), sw.astype(np.float32), normed


def kernel_matrix(row_ptr, col, vd, v):
    row_ptr = np.asarray(row_ptr, np.int64)
    n = len(row_ptr) - 1
    src = np.repeat(np.arange(n), np.diff(row_ptr))
    A = sp.csr_matrix((v.astype(np.float64), (src, np.asarray(col, np.int64))), shape=(n, n))
    return A + A.T + sp.diags(vd.astype(np.float64))


def set_data_box(y, box=10.0):
    """embedder.rs:1376-1408."""
    y = np.array(y, np.float64)
    y -= y.mean(axis=0)
    mm = np.abs(y).max() / (box / 2.0)
    return (y / mm).astype(np.float32)


def dmap_layout(row_ptr, col, dist, asked_dim=2, t=5.0, gnbn=12, alfa=0.5, beta=-0.1, rank=20, box=10.0, boxed=True):
    """The initial layout of embedder.rs:308-345 with an exact eigen-decomposition in place of the randomized SVD.
    Returns (layout[n, asked_dim], eigenvalues[rank] (descending), U[n, rank])."""
    vd, v, sw, normed = sym_kernel(row_ptr, col, dist, gnbn, alfa, beta)
    S = kernel_matrix(row_ptr, col, vd, v)
    n = S.shape[0]
    if n <= 3000:
        lam, U = np.linalg.eigh(S.toarray())
    else:
        import scipy.sparse.linalg as spl
        lam, U = spl.eigsh(S, k=rank, which="LA")
    # the reference takes singular values: |lambda| in decreasing order
    order = np.argsort(-np.abs(lam))[:rank]
    lam, U = np.abs(lam[order]), U[:, order]
    nl = lam / lam[0]
    weight = normed.astype(np.float64) * np.sqrt(sw.astype(np.float64) / sw.astype(np.float64).mean())   # :1221-1224
    real_dim = min(asked_dim, U.shape[1] - 1)
    y = np.zeros((n, real_dim))
    for j in range(real_dim):
        y[:, j] = np.clip(nl[j + 1] ** t * U[:, j + 1] / weight, -10.0, 10.0)
    y = y.astype(np.float32)
    return (set_data_box(y, box) if boxed else y), lam, U


def subspace_svd(S, rank=20, nbiter=5, seed=0, omega=None):
    """tools/svdapprox.rs:343-410 (subspace_iteration_csr: Gaussian test matrix, QR after every product) followed by
    :721-801 (direct_svd: B = Q^T S, SVD of B, U = Q U_B).  Returns (sigma[rank], U[n, rank])."""
    n = S.shape[0]
    rng = np.random.default_rng(seed)
    y = S @ (rng.standard_normal((n, rank)) if omega is None else np.asarray(omega, np.float64))
    q, _ = np.linalg.qr(y)
    for _ in range(1, nbiter):
        q, _ = np.linalg.qr(S.T @ q)
        q, _ = np.linalg.qr(S @ q)
    b = (S.T @ q).T                                        # Q^T S, rank x n
    ub, s, _ = np.linalg.svd(b, full_matrices=False)
    return s, q @ ub


def dmap_layout_randomized(row_ptr, col, dist, asked_dim=2, t=5.0, gnbn=12, seed=0, box=10.0, omega=None):
    """dmap_layout with the reference's randomized SVD instead of the exact one."""
    vd, v, sw, normed = sym_kernel(row_ptr, col, dist, gnbn)
    S = kernel_matrix(row_ptr, col, vd, v)
    lam, U = subspace_svd(S, seed=seed, omega=omega)
    nl = lam / lam[0]
    weight = normed.astype(np.float64) * np.sqrt(sw.astype(np.float64) / sw.astype(np.float64).mean())
    y = np.stack([np.clip(nl[j + 1] ** t * U[:, j + 1] / weight, -10.0, 10.0) for j in range(asked_dim)], axis=1)
    return set_data_box(y.astype(np.float32), box), lam, U
