"""ctypes binder for oracle/liboracle.so plus an independent numpy restatement.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU arm.
The product package (annembed_b200/) never imports this module.

PARITY UNPINNED (see annembed_oracle.c): the Rust reference cannot be run here and its tests hold
no golden vectors for this path.  The C restatement and the numpy restatement below were written
independently from the cited reference lines and are checked against each other and against
hand-derived vectors in tests/test_oracle.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PROBA_MIN = np.float32(1.0e-4)  # embedder.rs:50


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "annembed_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_cross_entropy.restype = C.c_double
        _LIB.oracle_optimize.restype = C.c_int64
        _LIB.oracle_optimize_reference_layout.restype = C.c_int64
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _graph(row_ptr, col):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
    col = np.ascontiguousarray(col, dtype=np.uint32)
    return row_ptr, col


# ------------------------------------------------------------------ C restatement wrappers
def edge_weights(row_ptr, col, dist, scale_rho=1.0, beta=1.0):
    """tools/kdumap.rs:26-235.  Returns (scale[N], p[E]); raises on an empty row (kdumap.rs:75-85)."""
    row_ptr, col = _graph(row_ptr, col)
    dist = np.ascontiguousarray(dist, dtype=np.float32)
    n = len(row_ptr) - 1
    scale = np.empty(n, np.float32)
    p = np.empty(len(col), np.float32)
    rc = lib().oracle_edge_weights(C.c_uint64(n), _p(row_ptr, C.c_uint64), _p(col, C.c_uint32),
                                   _p(dist, C.c_float), C.c_float(scale_rho), C.c_float(beta),
                                   _p(scale, C.c_float), _p(p, C.c_float))
    if rc != 0:
        raise ValueError(f"node {-rc - 1} has no neighbour")
    return scale, p


def perplexity(row_ptr, p):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
    p = np.ascontiguousarray(p, dtype=np.float32)
    out = np.empty(len(row_ptr) - 1, np.float32)
    lib().oracle_perplexity(C.c_uint64(len(out)), _p(row_ptr, C.c_uint64), _p(p, C.c_float), _p(out, C.c_float))
    return out


def embedded_scales(scale, f64_sum=False):
    """embedder.rs:1356-1373."""
    scale = np.ascontiguousarray(scale, dtype=np.float32)
    out = np.empty_like(scale)
    lib().oracle_embedded_scales(C.c_uint64(len(scale)), _p(scale, C.c_float), C.c_int(int(f64_sum)), _p(out, C.c_float))
    return out


def step_fixed(row_ptr, col, p, emb_scale, y, b, grad_step, edge_idx, neg):
    """embedder.rs:1167-1302 applied to an explicit (edge, 5 negatives) list, in order.  Returns new y."""
    row_ptr, col = _graph(row_ptr, col)
    p = np.ascontiguousarray(p, np.float32)
    emb_scale = np.ascontiguousarray(emb_scale, np.float32)
    y = np.array(y, dtype=np.float32, order="C", copy=True)
    edge_idx = np.ascontiguousarray(edge_idx, np.uint64)
    neg = np.ascontiguousarray(neg, np.uint32).reshape(-1, 5)
    assert len(neg) == len(edge_idx)
    n, d = y.shape
    rc = lib().oracle_step_fixed(C.c_uint64(n), C.c_uint32(d), _p(row_ptr, C.c_uint64), _p(col, C.c_uint32),
                                 _p(p, C.c_float), _p(emb_scale, C.c_float), _p(y, C.c_float),
                                 C.c_double(b), C.c_double(grad_step), C.c_uint64(len(edge_idx)),
                                 _p(edge_idx, C.c_uint64), _p(neg, C.c_uint32))
    assert rc == 0
    return y


def cross_entropy(row_ptr, col, p, emb_scale, y, b=1.0):
    """embedder.rs:1127-1163,1322-1345."""
    row_ptr, col = _graph(row_ptr, col)
    p = np.ascontiguousarray(p, np.float32)
    emb_scale = np.ascontiguousarray(emb_scale, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    n, d = y.shape
    return float(lib().oracle_cross_entropy(C.c_uint64(n), C.c_uint32(d), _p(row_ptr, C.c_uint64),
                                            _p(col, C.c_uint32), _p(p, C.c_float), _p(emb_scale, C.c_float),
                                            _p(y, C.c_float), C.c_double(b)))


def hubness_weights(row_ptr, col):
    """fromhnsw/hubness.rs:39-76 + the clamp of embedder.rs:826-833."""
    row_ptr, col = _graph(row_ptr, col)
    n = len(row_ptr) - 1
    w = np.empty(n, np.float32)
    lib().oracle_hubness_weights(C.c_uint64(n), _p(row_ptr, C.c_uint64), _p(col, C.c_uint32), _p(w, C.c_float))
    return w


def optimize(row_ptr, col, p, emb_scale, y0, b=1.0, grad_step=2.0, nb_sampling_by_edge=10, nb_grad_batch=20,
             neg_w=None, seed=0, first_batch=1, n_batches=None, sample_fraction=1.0, n_threads=0, timing=False,
             reference_layout=False):
    """embedder.rs:794-904 Hogwild loop.  Returns (y, positive_samples_processed[, loop_seconds]).
    reference_layout=True runs the same loop on the reference's data layout (per-node heap rows behind Arc + RwLock,
    heap copies per access: oracle_optimize_reference_layout) -- the timing twin of SURVEY.md 8(d)."""
    row_ptr, col = _graph(row_ptr, col)
    p = np.ascontiguousarray(p, np.float32)
    emb_scale = np.ascontiguousarray(emb_scale, np.float32)
    y = np.array(y0, dtype=np.float32, order="C", copy=True)
    n, d = y.shape
    if n_batches is None:
        n_batches = nb_grad_batch
    secs = C.c_double(0.0)
    negp = None
    if neg_w is not None:
        neg_w = np.ascontiguousarray(neg_w, np.float32)
        negp = _p(neg_w, C.c_float)
    fn = lib().oracle_optimize_reference_layout if reference_layout else lib().oracle_optimize
    done = fn(C.c_uint64(n), C.c_uint32(d), _p(row_ptr, C.c_uint64), _p(col, C.c_uint32),
              _p(p, C.c_float), _p(emb_scale, C.c_float), _p(y, C.c_float), C.c_double(b),
              C.c_double(grad_step), C.c_uint32(nb_sampling_by_edge), C.c_uint32(nb_grad_batch),
              negp, C.c_uint64(seed), C.c_uint32(first_batch), C.c_uint32(n_batches),
              C.c_double(sample_fraction), C.c_int(n_threads), C.byref(secs))
    if done < 0:
        raise RuntimeError(f"oracle_optimize failed: {done}")
    if timing:
        return y, int(done), secs.value
    return y, int(done)


def transformed_kgraph(row_ptr, col, y):
    """embedder.rs:478-522 (running minimum, then sorted)."""
    row_ptr, col = _graph(row_ptr, col)
    y = np.ascontiguousarray(y, np.float32)
    n, d = y.shape
    out = np.empty(len(col), np.float32)
    lib().oracle_transformed_kgraph(C.c_uint64(n), C.c_uint32(d), _p(row_ptr, C.c_uint64), _p(col, C.c_uint32),
                                    _p(y, C.c_float), _p(out, C.c_float))
    return out


def dichotomy_reference_tests():
    """The reference's two known-answer tests, tools/dichotomy.rs:75-91."""
    r1, r2 = C.c_float(), C.c_float()
    rc1 = lib().oracle_dichotomy_test_inc(C.byref(r1))
    rc2 = lib().oracle_dichotomy_test_dec(C.byref(r2))
    return (rc1, r1.value), (rc2, r2.value)


def scale_from_umap(d, norm):
    """embedder.rs:760-783 (dead code in the reference)."""
    d = np.ascontiguousarray(d, np.float32)
    w = np.empty_like(d)
    s = C.c_float()
    rc = lib().oracle_scale_from_umap(_p(d, C.c_float), C.c_uint64(len(d)), C.c_float(norm), C.byref(s), _p(w, C.c_float))
    return rc, s.value, w


def num_threads():
    return int(lib().oracle_num_threads())


# ------------------------------------------------------------------ independent numpy restatement
def np_edge_weights(row_ptr, col, dist, scale_rho=1.0, beta=1.0):
    """Second, independent restatement of tools/kdumap.rs:132-235 (pure Python loops: small inputs)."""
    f32 = np.float32
    n = len(row_ptr) - 1
    scale = np.zeros(n, f32)
    p = np.zeros(len(col), f32)
    for i in range(n):
        lo, hi = int(row_ptr[i]), int(row_ptr[i + 1])
        k = hi - lo
        if k == 0:
            raise ValueError(f"node {i} has no neighbour")
        d = dist[lo:hi].astype(f32)
        acc = f32(0)
        for m in range(lo, hi):
            acc = f32(acc + f32(dist[int(row_ptr[int(col[m])])]))
        acc = f32(acc + d[0])
        s = f32(f32(scale_rho) * f32(acc / f32(k + 1)))
        scale[i] = s
        pos = np.nonzero(d > 0)[0]
        if len(pos) and d[pos[-1]] > d[0]:
            w = np.zeros(k, f32)
            with np.errstate(all="ignore"):
                for m in range(k):
                    a = f32(max(f32(d[m] - d[0]), f32(0))) / s
                    v = np.exp(-np.power(f32(a), f32(beta)), dtype=f32)
                    w[m] = PROBA_MIN if (np.isnan(v) or v < PROBA_MIN) else v
            tot = f32(0)
            for m in range(k):
                tot = f32(tot + w[m])
            p[lo:hi] = w / tot
        else:
            p[lo:hi] = f32(1.0) / f32(k)
    return scale, p


def np_sgd_sample(y, i, j, p, scale, b, grad_step, negs):
    """Second restatement of embedder.rs:1167-1302 for one sample; y modified in place (f32 rows)."""
    f32 = np.float32
    yi = y[i].copy()
    yj = y[j].copy()
    g = np.zeros_like(yi)
    s2 = float(scale) * float(scale)

    def coeff(u):
        if b != 1.0:
            return 2.0 * b * (1.0 / (1.0 + u ** b)) * u ** (b - 1.0) / s2
        return 2.0 * b * (1.0 / (1.0 + u)) / s2

    dsum = f32(0)
    for c in range(len(yi)):
        dsum = f32(dsum + f32(f32(yi[c] - yj[c]) * f32(yi[c] - yj[c])))
    u = float(dsum) / s2
    if u > 0.0:
        rep = 1.0 / max(u * u, float(f32(1.0) / PROBA_MIN))
        cij = max(grad_step * coeff(u) * (-p + (1.0 - p) * rep), -0.49)
        g = ((yj - yi) * f32(cij)).astype(f32)
    yi = (yi - g).astype(f32)
    yj = (yj + g).astype(f32)
    y[j] = yj
    for k in negs:
        yk = y[k].copy()
        dk = f32(0)
        for c in range(len(yi)):
            dk = f32(dk + f32(f32(yi[c] - yk[c]) * f32(yi[c] - yk[c])))
        dik = float(dk)
        if dik > 0.0:
            uk = dik / s2
            rep = 1.0 / max(uk * uk, 1.0 / 16.0)
            cik = min(grad_step * coeff(uk) * rep, 2.0)
            g = ((yk - yi) * f32(cik)).astype(f32)
        yi = (yi - g).astype(f32)
    y[i] = yi


def np_cross_entropy(row_ptr, col, p, emb_scale, y, b=1.0):
    """Second restatement of embedder.rs:1127-1163,1322-1345."""
    f32 = np.float32
    ce = 0.0
    n = len(row_ptr) - 1
    for i in range(n):
        s = float(emb_scale[i])
        for m in range(int(row_ptr[i]), int(row_ptr[i + 1])):
            diff = (y[i] - y[int(col[m])]).astype(f32)
            ds = f32(0)
            for c in range(len(diff)):
                ds = f32(ds + f32(diff[c] * diff[c]))
            x = (float(ds) / (s * s)) ** b
            w = f32(1.0 / (1.0 + x))
            if not (w < f32(1)):
                w = f32(1) - np.finfo(f32).eps
            w = float(w)
            pe = float(p[m])
            if w > 0:
                ce += -pe * np.log(w)
            if w < 1:
                ce += -(1.0 - pe) * np.log(1.0 - w)
    return ce
