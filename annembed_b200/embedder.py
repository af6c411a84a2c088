"""Embedder -- host-side mirror of /root/reference/src/embedder.rs `Embedder` for the hot path only.

Same method names, argument meaning and error behaviour as the Rust struct (embedder.rs:84-453), with the pair
`to_proba_edges` + `entropy_optimize` (embedder.rs:351-356) replaced by calls into the C ABI
(include/annembed_cuda.h).  Also here: the device diffusion-map initial layout (dmap_init,
embedder.rs:308-345), the hierarchical `from_hkgraph` second step (embedder.rs:194-295) and the device quality estimate
(embedder.rs:620-753).  There is no CPU fallback: every method fails if the CUDA library or a device is missing.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import AnnembedCudaError, Params, Quality, Stats, ptr
from .embedparams import EmbedderParams
from .kgraph import KGraph


def _c_params(p: EmbedderParams) -> Params:
    return Params(asked_dim=p.asked_dim, dmap_init=int(p.dmap_init), beta=p.beta, b=p.b, scale_rho=p.scale_rho,
                  grad_step=p.grad_step, nb_sampling_by_edge=p.nb_sampling_by_edge, nb_grad_batch=p.nb_grad_batch,
                  grad_factor=p.grad_factor, hierarchy_layer=p.hierarchy_layer,
                  hubness_weighting=int(p.hubness_weighting), mini_epochs_per_batch=p.mini_epochs_per_batch,
                  seed=p.seed, flags=p.flags, cell_substeps=p.cell_substeps)


class CudaContext:
    """Thin RAII wrapper of annembed_cuda_ctx; every method maps 1:1 to a C-ABI call and raises on status != 0."""

    def __init__(self, params: EmbedderParams, device: int = 0):
        self.lib = _lib.load()
        self.params = params
        self.h = C.c_void_p()
        cp = _c_params(params)
        st = self.lib.annembed_cuda_create(C.byref(self.h), C.byref(cp), device)
        if st != 0:
            msg = self.lib.annembed_cuda_last_error(None)
            self.h = C.c_void_p()
            raise AnnembedCudaError(st, msg.decode() if msg else "")
        self.n = 0
        self.E = 0

    def _ck(self, st: int):
        if st != 0:
            msg = self.lib.annembed_cuda_last_error(self.h)
            raise AnnembedCudaError(st, msg.decode() if msg else "")

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.annembed_cuda_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- multi GPU
    def unique_id(self) -> np.ndarray:
        out = np.zeros(128, np.uint8)
        st = self.lib.annembed_cuda_comm_unique_id(ptr(out, C.c_uint8))
        if st != 0:
            raise AnnembedCudaError(st, (self.lib.annembed_cuda_last_error(None) or b"").decode())
        return out

    def comm_init(self, rank: int, nranks: int, unique_id: np.ndarray | None):
        uid = None if unique_id is None else np.ascontiguousarray(unique_id, np.uint8)
        self._ck(self.lib.annembed_cuda_comm_init(self.h, rank, nranks, ptr(uid, C.c_uint8)))

    def export_layout(self) -> np.ndarray:
        out = np.zeros(128, np.uint8)
        self._ck(self.lib.annembed_cuda_comm_export_layout(self.h, ptr(out, C.c_uint8)))
        return out

    def import_layouts(self, all_handles: np.ndarray):
        h = np.ascontiguousarray(all_handles, np.uint8).reshape(-1)
        self._ck(self.lib.annembed_cuda_comm_import_layouts(self.h, ptr(h, C.c_uint8)))

    # --- graph / weights
    def set_graph_csr(self, row_ptr, col, dist):
        row_ptr = np.ascontiguousarray(row_ptr, np.uint64)
        col = np.ascontiguousarray(col, np.uint32)
        dist = np.ascontiguousarray(dist, np.float32)
        n = len(row_ptr) - 1
        if n < 0 or len(col) != len(dist) or (n >= 0 and int(row_ptr[-1]) != len(col)):
            raise ValueError("inconsistent CSR arrays (len(col) == len(dist) == row_ptr[-1] required)")
        self._ck(self.lib.annembed_cuda_set_graph_csr(self.h, n, ptr(row_ptr, C.c_uint64), ptr(col, C.c_uint32),
                                                      ptr(dist, C.c_float)))
        self.n, self.E = n, len(col)

    def edge_weights(self, want_outputs: bool = True):
        if not want_outputs:
            self._ck(self.lib.annembed_cuda_edge_weights(self.h, None, None))
            return None, None
        scale = np.empty(self.n, np.float32)
        p = np.empty(self.E, np.float32)
        self._ck(self.lib.annembed_cuda_edge_weights(self.h, ptr(scale, C.c_float), ptr(p, C.c_float)))
        return scale, p

    def edge_weights_umap(self, norm: float):
        scale = np.empty(self.n, np.float32)
        w = np.empty(self.E, np.float32)
        status = np.empty(self.n, np.uint8)
        self._ck(self.lib.annembed_cuda_edge_weights_umap(self.h, norm, ptr(scale, C.c_float), ptr(w, C.c_float),
                                                          ptr(status, C.c_uint8)))
        return scale, w, status

    def set_edge_weights(self, scale, proba):
        scale = np.ascontiguousarray(scale, np.float32)
        proba = np.ascontiguousarray(proba, np.float32)
        assert len(scale) == self.n and len(proba) == self.E
        self._ck(self.lib.annembed_cuda_set_edge_weights(self.h, ptr(scale, C.c_float), ptr(proba, C.c_float)))

    def get_perplexity(self):
        out = np.empty(self.n, np.float32)
        self._ck(self.lib.annembed_cuda_get_perplexity(self.h, ptr(out, C.c_float)))
        return out

    def set_neg_weights(self, w):
        if w is not None:
            w = np.ascontiguousarray(w, np.float32)
            assert len(w) == self.n
        self._ck(self.lib.annembed_cuda_set_neg_weights(self.h, ptr(w, C.c_float)))

    def get_hubness_counts(self):
        out = np.empty(self.n, np.uint32)
        self._ck(self.lib.annembed_cuda_get_hubness_counts(self.h, ptr(out, C.c_uint32)))
        return out

    # --- layout
    def set_embedding(self, y):
        y = np.ascontiguousarray(y, np.float32)
        if y.shape != (self.n, self.params.asked_dim):
            raise ValueError(f"initial embedding must be ({self.n}, {self.params.asked_dim}), got {y.shape}")
        self._ck(self.lib.annembed_cuda_set_embedding(self.h, ptr(y, C.c_float)))

    def set_embedding_from_projection(self, first, proj_node, proj_dist, median_dist: float):
        first = np.ascontiguousarray(first, np.float32)
        proj_node = np.ascontiguousarray(proj_node, np.uint32)
        proj_dist = np.ascontiguousarray(proj_dist, np.float32)
        assert first.shape[1] == self.params.asked_dim and len(proj_node) == self.n == len(proj_dist)
        self._ck(self.lib.annembed_cuda_set_embedding_from_projection(self.h, first.shape[0], ptr(first, C.c_float),
                                                                      ptr(proj_node, C.c_uint32), ptr(proj_dist, C.c_float),
                                                                      median_dist))

    def dmap_init(self, gnbn: int = 12, diffusion_time: float = 5.0, want_output: bool = True):
        """≙ the dmap_init branch of one_step_embed (embedder.rs:308-345) on the device: computes the diffusion-map
        layout of the loaded graph, installs it as the initial embedding and returns it (n x asked_dim)."""
        out = np.empty((self.n, self.params.asked_dim), np.float32) if want_output else None
        self._ck(self.lib.annembed_cuda_dmap_init(self.h, int(gnbn), float(diffusion_time), ptr(out, C.c_float)))
        return out

    def dmap_kernel(self, gnbn: int = 12):
        """(diag[n], val[E], sw[n], normed_scale[n]) of the symmetric normalised kernel (see annembed_cuda_dmap_kernel)."""
        diag = np.empty(self.n, np.float32); val = np.empty(self.E, np.float32)
        sw = np.empty(self.n, np.float32); normed = np.empty(self.n, np.float32)
        self._ck(self.lib.annembed_cuda_dmap_kernel(self.h, int(gnbn), ptr(diag, C.c_float), ptr(val, C.c_float),
                                                    ptr(sw, C.c_float), ptr(normed, C.c_float)))
        return diag, val, sw, normed

    def dmap_set_test_matrix(self, omega):
        if omega is None:
            self._ck(self.lib.annembed_cuda_dmap_set_test_matrix(self.h, None, 0))
            return
        omega = np.ascontiguousarray(omega, np.float32)
        assert omega.shape == (self.n, 20)
        self._ck(self.lib.annembed_cuda_dmap_set_test_matrix(self.h, ptr(omega, C.c_float), self.n))

    def dmap_singular_values(self, count: int = 20) -> np.ndarray:
        out = np.zeros(count, np.float64)
        self._ck(self.lib.annembed_cuda_dmap_singular_values(self.h, ptr(out, C.c_double), count))
        return out

    def reset_embedding(self):
        self._ck(self.lib.annembed_cuda_reset_embedding(self.h))

    def get_embedded_scales(self):
        out = np.empty(self.n, np.float32)
        self._ck(self.lib.annembed_cuda_get_embedded_scales(self.h, ptr(out, C.c_float)))
        return out

    def step_fixed(self, edge_idx, neg_idx, grad_step: float):
        edge_idx = np.ascontiguousarray(edge_idx, np.uint64)
        neg_idx = np.ascontiguousarray(neg_idx, np.uint32).reshape(-1)
        assert len(neg_idx) == 5 * len(edge_idx)
        self._ck(self.lib.annembed_cuda_step_fixed(self.h, len(edge_idx), ptr(edge_idx, C.c_uint64),
                                                   ptr(neg_idx, C.c_uint32), grad_step))

    def optimize(self, want_ce: bool = True):
        if want_ce:
            a, b = C.c_double(), C.c_double()
            self._ck(self.lib.annembed_cuda_optimize(self.h, C.byref(a), C.byref(b)))
            return a.value, b.value
        self._ck(self.lib.annembed_cuda_optimize(self.h, None, None))
        return None, None

    def optimize_batches(self, first_batch: int, n_batches: int):
        self._ck(self.lib.annembed_cuda_optimize_batches(self.h, first_batch, n_batches))

    def cross_entropy(self) -> float:
        v = C.c_double()
        self._ck(self.lib.annembed_cuda_cross_entropy(self.h, C.byref(v)))
        return v.value

    def get_embedding(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = _result_buffer((self.n, self.params.asked_dim))
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == self.n * self.params.asked_dim
        self._ck(self.lib.annembed_cuda_get_embedding(self.h, ptr(out, C.c_float)))
        return out

    def quality_estimate(self, nbng: int, want_arrays: bool = False) -> dict:
        q = Quality()
        radius = np.empty(self.n, np.float32) if want_arrays else None
        first = np.empty(self.n, np.float32) if want_arrays else None
        nratio = np.empty(self.n, np.float32) if want_arrays else None
        self._ck(self.lib.annembed_cuda_quality_estimate(self.h, nbng, C.byref(q), ptr(radius, C.c_float),
                                                         ptr(first, C.c_float), ptr(nratio, C.c_float)))
        out = {"nb_without_match": int(q.nb_without_match), "mean_nbmatch": q.mean_nbmatch,
               "knn_preservation": q.knn_preservation, "mean_ratio": q.mean_ratio,
               "radius_quantiles": list(q.radius_quantiles), "ratio_quantiles": list(q.ratio_quantiles),
               "median_ratio": q.ratio_quantiles[2]}
        if want_arrays:
            out.update(radius=radius, first_dist=first, ratio_by_node=nratio)
        return out

    def get_stats(self) -> dict:
        s = Stats()
        self._ck(self.lib.annembed_cuda_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        self._ck(self.lib.annembed_cuda_reset_stats(self.h))

    def debug_draws(self, epoch: int, want_negs: bool = True):
        counts = np.empty(self.E, np.uint32)
        negs = np.empty((self.E, 5), np.uint32) if want_negs else None
        self._ck(self.lib.annembed_cuda_debug_draws(self.h, epoch, ptr(counts, C.c_uint32), ptr(negs, C.c_uint32)))
        return counts, negs


def _result_buffer(shape) -> np.ndarray:
    """Host array for a layout read back from the device: page-locked when torch can provide it (its caching host
    allocator recycles the block once the array is dropped), so that the copy runs at PCIe speed -- a pageable 88 MB
    destination took 21 ms at 11M nodes, 2 ms pinned.  Plain numpy otherwise."""
    try:
        import torch
        if torch.cuda.is_available():
            return torch.empty(tuple(shape), dtype=torch.float32, pin_memory=True).numpy()
    except Exception:
        pass
    return np.empty(shape, np.float32)


class KGraphProjection:
    """≙ `KGraphProjection<F>` (fromhnsw/kgproj.rs:35-44) as flat arrays: a small graph on the upper HNSW layers, the
    whole (large) graph whose first `small_graph.get_nb_nodes()` nodes ARE the small graph's nodes (kgproj.rs:106-124),
    and for every other node its projection = nearest small-graph node and the distance to it (proj_data)."""

    def __init__(self, small_graph: KGraph, large_graph: KGraph, proj_node, proj_dist, layer: int = 1):
        self.layer = layer
        self.small_graph = small_graph
        self.large_graph = large_graph
        self.proj_node = np.ascontiguousarray(proj_node, np.uint32)
        self.proj_dist = np.ascontiguousarray(proj_dist, np.float32)
        n, ns = large_graph.get_nb_nodes(), small_graph.get_nb_nodes()
        if len(self.proj_node) != n or len(self.proj_dist) != n or ns > n:
            raise ValueError("projection arrays must have one entry per node of the large graph")

    def get_small_graph(self) -> KGraph: return self.small_graph          # kgproj.rs:388
    def get_large_graph(self) -> KGraph: return self.large_graph          # :393
    def get_projection_by_nodeidx(self, i: int): return int(self.proj_node[i]), float(self.proj_dist[i])   # :376

    def get_projection_distance_median(self) -> float:
        """≙ get_projection_distance_quant().query(0.5) (kgproj.rs:403-410; exact median instead of CKMS)."""
        ns = self.small_graph.get_nb_nodes()
        return float(np.median(self.proj_dist[ns:])) if len(self.proj_dist) > ns else 1.0


class EmbedError(RuntimeError):
    """≙ `Err(1)` of Embedder::embed (embedder.rs:183,366-369)."""


class Embedder:
    """≙ `Embedder<'a, F>` (embedder.rs:84-100) restricted to the one-step path (`Embedder::new`, :107)."""

    def __init__(self, kgraph: KGraph, parameters: EmbedderParams, initial_embedding: np.ndarray | None = None,
                 device: int = 0, comm: tuple | None = None, fused_exchange: bool = True,
                 context: CudaContext | None = None):
        self.kgraph = kgraph                     # borrowed, like &'a KGraph<F>
        self.parameters = parameters             # copied by value in the reference (EmbedderParams: Copy)
        self.initial_embedding = None if initial_embedding is None else np.ascontiguousarray(initial_embedding, np.float32)
        self.embedding = None
        self.hubness_counts = None
        self.cross_entropy = (None, None)
        self.device = device
        self.comm = comm                          # (rank, nranks, unique_id) or None
        self.fused_exchange = fused_exchange      # open the peers' layout buffers (asynchronous form on several ranks)
        self.context = context                    # optional long-lived device context (keeps its NCCL communicator and
                                                  # peer mappings across embeds); by default embed() creates and destroys one
        self.write_quality_csv = False            # the reference dumps first_dist.csv / continuity_ratio.csv (embedder.rs:729-743)
        self.stats = {}

    @classmethod
    def from_hkgraph(cls, graph_projection: KGraphProjection, parameters: EmbedderParams,
                     initial_embedding: np.ndarray | None = None, device: int = 0):
        """≙ Embedder::from_hkgraph (embedder.rs:120-133): two-step embedding.  `initial_embedding` is the initial
        layout of the SMALL graph (the reference computes a diffusion-map layout for it)."""
        e = cls(graph_projection.get_large_graph(), parameters, initial_embedding, device)
        e.hkgraph = graph_projection
        return e

    def h_embed(self) -> int:
        """≙ h_embed (embedder.rs:194-295)."""
        import dataclasses
        proj, p = self.hkgraph, self.parameters
        first_params = dataclasses.replace(p, nb_grad_batch=p.grad_factor * p.nb_grad_batch, grad_step=1.0,
                                           hierarchy_layer=0)                       # embedder.rs:203-209
        first = Embedder(proj.get_small_graph(), first_params, self.initial_embedding, self.device)
        first.embed()                                                               # :213
        first_embedding = first.get_embedded()
        self.first_step_stats = first.stats
        ctx = None
        try:
            large = proj.get_large_graph()
            ctx = CudaContext(p, self.device)
            ctx.set_graph_csr(*large.get_neighbours())
            ctx.edge_weights(want_outputs=False)                                    # :226-230
            if p.hubness_weighting:
                counts = ctx.get_hubness_counts()
                self.hubness_counts = counts
                ctx.set_neg_weights(np.clip(counts.astype(np.float32), 1.0, float(len(counts))))
            ctx.set_embedding_from_projection(first_embedding, proj.proj_node, proj.proj_dist,
                                              proj.get_projection_distance_median())   # :245-269
            self.initial_embedding = ctx.get_embedding()
            self.cross_entropy = ctx.optimize(want_ce=True)                         # :275
            self.embedding = ctx.get_embedding()
            self.stats = ctx.get_stats()
        except AnnembedCudaError as e:
            raise EmbedError(str(e)) from e
        finally:
            if ctx is not None:
                ctx.close()
        return 1

    # --- parameter getters, embedder.rs:135-153
    def get_asked_dimension(self) -> int: return self.parameters.asked_dim
    def get_scale_rho(self) -> float: return self.parameters.scale_rho
    def get_b(self) -> float: return self.parameters.b
    def get_grad_step(self) -> float: return self.parameters.grad_step
    def get_nb_grad_batch(self) -> int: return self.parameters.nb_grad_batch
    def get_kgraph(self) -> KGraph: return self.kgraph
    def get_hubness(self): return self.hubness_counts
    def get_nb_nodes(self) -> int: return self.kgraph.get_nb_nodes()

    def _get_random_init(self, size: float) -> np.ndarray:
        """≙ get_random_init (embedder.rs:456-470): uniform in [-size/2, size/2]^d (seeded here)."""
        rng = np.random.Generator(np.random.PCG64(self.parameters.seed))
        n, d = self.get_nb_nodes(), self.parameters.asked_dim
        return rng.uniform(-size / 2, size / 2, size=(n, d)).astype(np.float32)

    def embed(self) -> int:
        """≙ embed -> one_step_embed (embedder.rs:183,298-371).  Returns 1 (`Ok(1)`); raises EmbedError (`Err(1)`)."""
        p = self.parameters
        if getattr(self, "hkgraph", None) is not None:
            return self.h_embed()                                                   # embedder.rs:186-190
        device_dmap = self.initial_embedding is None and p.dmap_init    # embedder.rs:308-345 on the device
        if device_dmap and self.get_nb_nodes() < 80:
            # the device range finder is the reference's rank-20 one (graphlaplace.rs:113): it needs 4 x 20 nodes.  The
            # reference switches to a full SVD below 500 nodes (graphlaplace.rs:100-108), which the device path does not
            # have: tiny graphs (the upper layers of a small hierarchy) start from the random layout of the
            # dmap_init = false branch instead.  (asked_dim > 19 stays an error: EmbedError / ANNEMBED_ERR_UNSUPPORTED.)
            import warnings
            warnings.warn("dmap_init needs >= 80 nodes on the device: using the random initial layout (embedder.rs:348)",
                          RuntimeWarning, stacklevel=2)
            device_dmap = False
        if self.initial_embedding is None and not device_dmap:
            self.initial_embedding = self._get_random_init(1.0)       # embedder.rs:348
        ctx = self.context
        own_ctx = ctx is None
        import time as _time
        lap = [_time.perf_counter()]
        self.host_timings_ms = tm = {}

        def mark(name):                                  # host wall time per phase (the calls are synchronous)
            lap.append(_time.perf_counter())
            tm[name] = tm.get(name, 0.0) + 1e3 * (lap[-1] - lap[-2])
        try:
            if own_ctx:
                ctx = CudaContext(p, self.device)
                if self.comm is not None:
                    ctx.comm_init(*self.comm)
            mark("create")
            row_ptr, col, dist = self.kgraph.get_neighbours()
            ctx.set_graph_csr(row_ptr, col, dist)
            mark("set_graph_csr")
            if own_ctx and self.comm is not None and self.comm[1] > 1 and self.fused_exchange:
                from .dist import exchange_layout_handles
                exchange_layout_handles(ctx, self.comm[0], self.comm[1])
                mark("exchange_layout_handles")
            ctx.edge_weights(want_outputs=False)                       # to_proba_edges, embedder.rs:351
            if p.hubness_weighting:                                    # embedder.rs:810-837
                counts = ctx.get_hubness_counts()
                self.hubness_counts = counts
                n = float(len(counts))
                ctx.set_neg_weights(np.clip(counts.astype(np.float32), 1.0, n))
            mark("edge_weights")
            if device_dmap:
                self.initial_embedding = ctx.dmap_init()
                mark("dmap_init")
            else:
                ctx.set_embedding(self.initial_embedding)
                mark("set_embedding")
            self.cross_entropy = ctx.optimize(want_ce=True)            # entropy_optimize, embedder.rs:356
            mark("optimize")
            self.embedding = ctx.get_embedding()
            mark("get_embedding")
            self.stats = ctx.get_stats()
        except AnnembedCudaError as e:
            raise EmbedError(str(e)) from e
        finally:
            if own_ctx and ctx is not None:
                ctx.close()
                mark("close")
        return 1

    def get_quality_estimate_from_edge_length(self, nbng: int) -> dict:
        """≙ embedder.rs:620-753 on the device (annembed_cuda_quality_estimate).  Returns the statistics the reference
        prints (it returns Some(0.) itself) and writes first_dist.csv / continuity_ratio.csv in the CWD like it does."""
        if self.embedding is None:
            raise RuntimeError("cannot ask for embedded quality before embedding (embedder.rs:629-632)")
        from .io import write_csv_labeled_array2
        ctx = CudaContext(self.parameters, self.device)
        try:
            ctx.set_graph_csr(*self.kgraph.get_neighbours())
            ctx.set_embedding(self.embedding)
            q = ctx.quality_estimate(nbng, want_arrays=True)
        except AnnembedCudaError as e:
            raise EmbedError(str(e)) from e
        finally:
            ctx.close()
        if self.write_quality_csv:
            re = self.get_embedded_reindexed()
            write_csv_labeled_array2("first_dist.csv", q["first_dist"], re)
            write_csv_labeled_array2("continuity_ratio.csv", q["ratio_by_node"], re)
        return q

    # --- results, embedder.rs:378-453
    def get_embedded(self):
        return self.embedding

    def get_embedded_reindexed(self) -> np.ndarray:
        """≙ embedder.rs:384-405: row i of the embedding goes to row DataId(i); DataIds must be 0..n."""
        if self.embedding is None:
            raise RuntimeError("get_embedded_reindexed called before embed (the reference panics, embedder.rs:385)")
        return self._reindex(self.embedding)

    def get_embedded_by_dataid(self, data_id: int) -> np.ndarray:
        return self.embedding[self.kgraph.get_idx_from_dataid(data_id)]

    def get_embedded_by_nodeid(self, node: int) -> np.ndarray:
        return self.embedding[node]

    def get_initial_embedding(self):
        return self.initial_embedding

    def get_initial_embedding_reindexed(self) -> np.ndarray:
        return self._reindex(self.initial_embedding)

    def _reindex(self, a: np.ndarray) -> np.ndarray:
        ids = self.kgraph.data_id.astype(np.int64)
        n = len(ids)
        if ids.min(initial=0) < 0 or ids.max(initial=-1) >= n or len(np.unique(ids)) != n:
            raise IndexError("DataIds must fill 0..n to reindex (embedder.rs:397-403 indexes out of bounds otherwise)")
        out = np.zeros_like(a)
        out[ids] = a
        return out
