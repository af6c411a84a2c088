"""ctypes access to tests/hostsim/libhostsim.so (TEST TOOLING: host build of the product's device code)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hostsim")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
        _LIB = C.CDLL(os.path.join(_HERE, "libhostsim.so"))
        _LIB.hostsim_optimize.restype = C.c_int64
        _LIB.hostsim_optimize_cells.restype = C.c_int64
    return _LIB


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def philox2(ctr, key):
    ctr = np.asarray(ctr, np.uint32); out = np.zeros(2, np.uint32)
    lib().hostsim_philox2(_p(ctr, C.c_uint32), C.c_uint32(key), _p(out, C.c_uint32))
    return out


def alias_tables(w, G):
    """(sector table [nsec, 8] u32, T1 [nl, 2] u32, T2 [nl, 2 G] u32) of annembed_b200/csrc/alias_tables.hpp, identity numbering."""
    w = np.ascontiguousarray(w, np.float32)
    n = len(w)
    nsec, nl = (n + 3) // 4, (n + G - 1) // G
    sec = np.zeros((nsec, 8), np.uint32); t1 = np.zeros((nl, 2), np.uint32); t2 = np.zeros((nl, 2 * G), np.uint32)
    lib().hostsim_alias_tables(C.c_uint64(n), C.c_uint32(G), _p(w, C.c_float), _p(sec, C.c_uint32), _p(t1, C.c_uint32), _p(t2, C.c_uint32))
    return sec, t1, t2


def node_alias_table(w):
    """Node-level alias table [n, 2] u32 = {bits(prob), alias} (annembed_b200/csrc/alias_tables.hpp)."""
    w = np.ascontiguousarray(w, np.float32)
    tab = np.zeros((len(w), 2), np.uint32)
    lib().hostsim_node_alias_table(C.c_uint64(len(w)), _p(w, C.c_float), _p(tab, C.c_uint32))
    return tab


def below(w, low8, n):
    w = np.ascontiguousarray(w, np.uint32); low8 = np.ascontiguousarray(low8, np.uint32)
    o32, o40, oa = (np.zeros(len(w), np.uint32) for _ in range(3))
    lib().hostsim_below(C.c_uint64(len(w)), _p(w, C.c_uint32), _p(low8, C.c_uint32), C.c_uint32(n), _p(o32, C.c_uint32),
                        _p(o40, C.c_uint32), _p(oa, C.c_uint32))
    return o32, o40, oa


def node_uniforms(node0, count, epoch, seed):
    out = np.zeros(count, np.float32)
    lib().hostsim_node_uniforms(C.c_uint32(node0), C.c_uint32(count), C.c_uint32(epoch), C.c_uint64(seed), _p(out, C.c_float))
    return out


def philox(ctr, key):
    ctr = np.asarray(ctr, np.uint32); key = np.asarray(key, np.uint32); out = np.zeros(4, np.uint32)
    lib().hostsim_philox(_p(ctr, C.c_uint32), _p(key, C.c_uint32), _p(out, C.c_uint32))
    return out


def optimize(row_ptr, col, p, emb_scale, y0, b=1.0, grad_step=2.0, nbs=10, nb_batch=20, M=10, seed=1, neg_alias=None,
             first_batch=1, n_batches=None):
    row_ptr = np.ascontiguousarray(row_ptr, np.uint64); col = np.ascontiguousarray(col, np.uint32)
    p = np.ascontiguousarray(p, np.float32); emb_scale = np.ascontiguousarray(emb_scale, np.float32)
    y = np.array(y0, np.float32, order="C", copy=True)
    n, d = y.shape
    if n_batches is None:
        n_batches = nb_batch
    done = lib().hostsim_optimize(C.c_uint64(n), C.c_uint32(d), _p(row_ptr, C.c_uint64), _p(col, C.c_uint32), _p(p, C.c_float),
                                  _p(emb_scale, C.c_float), _p(y, C.c_float), C.c_double(b), C.c_double(grad_step),
                                  C.c_uint32(nbs), C.c_uint32(nb_batch), C.c_uint32(M), C.c_uint64(seed),
                                  _p(neg_alias, C.c_uint32), C.c_uint32(first_batch), C.c_uint32(n_batches))
    return y, int(done)


def optimize_cells(row_ptr, col, p, emb_scale, y0, b=1.0, grad_step=2.0, nbs=10, nb_batch=20, M=10, seed=1, neg_alias=None,
                   first_batch=1, n_batches=None, cell_nodes=4096, substeps=1):
    """Host replay of the cell-resident epoch kernel (identity numbering, fixed grid of cells, `substeps` per launch)."""
    row_ptr = np.ascontiguousarray(row_ptr, np.uint64); col = np.ascontiguousarray(col, np.uint32)
    p = np.ascontiguousarray(p, np.float32); emb_scale = np.ascontiguousarray(emb_scale, np.float32)
    y = np.array(y0, np.float32, order="C", copy=True)
    n, d = y.shape
    if n_batches is None:
        n_batches = nb_batch
    done = lib().hostsim_optimize_cells(C.c_uint64(n), C.c_uint32(d), _p(row_ptr, C.c_uint64), _p(col, C.c_uint32), _p(p, C.c_float),
                                        _p(emb_scale, C.c_float), _p(y, C.c_float), C.c_double(b), C.c_double(grad_step),
                                        C.c_uint32(nbs), C.c_uint32(nb_batch), C.c_uint32(M), C.c_uint64(seed),
                                        _p(neg_alias, C.c_uint32), C.c_uint32(first_batch), C.c_uint32(n_batches),
                                        C.c_uint32(cell_nodes), C.c_uint32(substeps))
    assert done >= 0
    return y, int(done)


def draws(row_ptr, col, p, nbs, M, seed, epoch, neg_alias=None, want_negs=True):
    row_ptr = np.ascontiguousarray(row_ptr, np.uint64); col = np.ascontiguousarray(col, np.uint32)
    p = np.ascontiguousarray(p, np.float32)
    n = len(row_ptr) - 1
    counts = np.zeros(len(col), np.uint32)
    negs = np.zeros((len(col), 5), np.uint32) if want_negs else None
    lib().hostsim_draws(C.c_uint64(n), _p(row_ptr, C.c_uint64), _p(col, C.c_uint32), _p(p, C.c_float), C.c_uint32(nbs),
                        C.c_uint32(M), C.c_uint64(seed), C.c_uint32(epoch), _p(neg_alias, C.c_uint32),
                        _p(counts, C.c_uint32), _p(negs, C.c_uint32))
    return counts, negs


def step_fixed(row_ptr, col, p, emb_scale, y0, b, grad_step, edge_idx, negs):
    row_ptr = np.ascontiguousarray(row_ptr, np.uint64); col = np.ascontiguousarray(col, np.uint32)
    p = np.ascontiguousarray(p, np.float32); emb_scale = np.ascontiguousarray(emb_scale, np.float32)
    y = np.array(y0, np.float32, order="C", copy=True)
    edge_idx = np.ascontiguousarray(edge_idx, np.uint64); negs = np.ascontiguousarray(negs, np.uint32)
    n, d = y.shape
    lib().hostsim_step_fixed(C.c_uint64(n), C.c_uint32(d), _p(row_ptr, C.c_uint64), _p(col, C.c_uint32), _p(p, C.c_float),
                             _p(emb_scale, C.c_float), _p(y, C.c_float), C.c_double(b), C.c_double(grad_step),
                             C.c_uint64(len(edge_idx)), _p(edge_idx, C.c_uint64), _p(negs, C.c_uint32))
    return y
