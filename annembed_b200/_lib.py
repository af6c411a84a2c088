"""ctypes binding of the C ABI declared in include/annembed_cuda.h.

There is no CPU fallback: if the shared library is missing or no CUDA device is present, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

u8p, u32p, u64p, f32p, f64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_uint64, C.c_float, C.c_double))


class Params(C.Structure):
    """annembed_cuda_params: mirror of EmbedderParams (embedparams.rs:76-103) + device knobs."""
    _fields_ = [
        ("asked_dim", C.c_uint32), ("dmap_init", C.c_uint32),
        ("beta", C.c_double), ("b", C.c_double), ("scale_rho", C.c_double), ("grad_step", C.c_double),
        ("nb_sampling_by_edge", C.c_uint32), ("nb_grad_batch", C.c_uint32), ("grad_factor", C.c_uint32),
        ("hierarchy_layer", C.c_uint32), ("hubness_weighting", C.c_uint32),
        ("mini_epochs_per_batch", C.c_uint32), ("seed", C.c_uint64), ("flags", C.c_uint32), ("cell_substeps", C.c_uint32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("edge_weights_ms", C.c_double), ("build_ms", C.c_double), ("optimize_ms", C.c_double),
        ("epoch_kernel_ms", C.c_double), ("exchange_ms", C.c_double), ("cross_entropy_ms", C.c_double),
        ("epoch_launches", C.c_uint64), ("kernel_launches", C.c_uint64), ("positive_samples", C.c_uint64),
        ("edge_updates", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("model_bytes", C.c_double),
        ("mini_epochs_per_batch", C.c_uint64), ("l2_persist_max_bytes", C.c_uint64), ("l2_window_max_bytes", C.c_uint64),
        ("n_cells", C.c_uint64), ("cell_nodes", C.c_uint64), ("cell_substeps", C.c_uint64), ("cross_cell_edges", C.c_uint64),
        ("cross_rank_edges", C.c_uint64), ("exchanges", C.c_uint64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Quality(C.Structure):
    _fields_ = [("nb_without_match", C.c_uint64), ("mean_nbmatch", C.c_double), ("knn_preservation", C.c_double),
                ("mean_ratio", C.c_double), ("radius_quantiles", C.c_double * 6), ("ratio_quantiles", C.c_double * 6)]


STATUS = {0: "OK", 1: "INVALID_ARG", 2: "CUDA", 3: "EMPTY_ROW", 4: "UNSORTED_ROW", 5: "STATE", 6: "UNSUPPORTED",
          7: "COMM", 8: "NO_NEGATIVE"}

# every symbol include/annembed_cuda.h declares: (name, restype, argtypes)
_ctx = C.c_void_p
SYMBOLS = [
    ("annembed_cuda_default_params", C.c_int, [C.POINTER(Params)]),
    ("annembed_cuda_create", C.c_int, [C.POINTER(_ctx), C.POINTER(Params), C.c_int]),
    ("annembed_cuda_destroy", C.c_int, [_ctx]),
    ("annembed_cuda_last_error", C.c_char_p, [_ctx]),
    ("annembed_cuda_comm_unique_id", C.c_int, [u8p]),
    ("annembed_cuda_comm_init", C.c_int, [_ctx, C.c_int, C.c_int, u8p]),
    ("annembed_cuda_comm_export_layout", C.c_int, [_ctx, u8p]),
    ("annembed_cuda_comm_import_layouts", C.c_int, [_ctx, u8p]),
    ("annembed_cuda_set_graph_csr", C.c_int, [_ctx, C.c_uint64, u64p, u32p, f32p]),
    ("annembed_cuda_edge_weights", C.c_int, [_ctx, f32p, f32p]),
    ("annembed_cuda_edge_weights_umap", C.c_int, [_ctx, C.c_float, f32p, f32p, u8p]),
    ("annembed_cuda_set_edge_weights", C.c_int, [_ctx, f32p, f32p]),
    ("annembed_cuda_get_perplexity", C.c_int, [_ctx, f32p]),
    ("annembed_cuda_set_neg_weights", C.c_int, [_ctx, f32p]),
    ("annembed_cuda_get_hubness_counts", C.c_int, [_ctx, u32p]),
    ("annembed_cuda_set_embedding", C.c_int, [_ctx, f32p]),
    ("annembed_cuda_reset_embedding", C.c_int, [_ctx]),
    ("annembed_cuda_dmap_init", C.c_int, [_ctx, C.c_uint32, C.c_float, f32p]),
    ("annembed_cuda_dmap_kernel", C.c_int, [_ctx, C.c_uint32, f32p, f32p, f32p, f32p]),
    ("annembed_cuda_dmap_set_test_matrix", C.c_int, [_ctx, f32p, C.c_uint64]),
    ("annembed_cuda_dmap_singular_values", C.c_int, [_ctx, f64p, C.c_uint32]),
    ("annembed_cuda_set_embedding_from_projection", C.c_int, [_ctx, C.c_uint64, f32p, u32p, f32p, C.c_float]),
    ("annembed_cuda_get_embedded_scales", C.c_int, [_ctx, f32p]),
    ("annembed_cuda_step_fixed", C.c_int, [_ctx, C.c_uint64, u64p, u32p, C.c_double]),
    ("annembed_cuda_optimize", C.c_int, [_ctx, f64p, f64p]),
    ("annembed_cuda_optimize_batches", C.c_int, [_ctx, C.c_uint32, C.c_uint32]),
    ("annembed_cuda_cross_entropy", C.c_int, [_ctx, f64p]),
    ("annembed_cuda_get_embedding", C.c_int, [_ctx, f32p]),
    ("annembed_cuda_quality_estimate", C.c_int, [_ctx, C.c_uint32, C.POINTER(Quality), f32p, f32p, f32p]),
    ("annembed_cuda_get_stats", C.c_int, [_ctx, C.POINTER(Stats)]),
    ("annembed_cuda_reset_stats", C.c_int, [_ctx]),
    ("annembed_cuda_debug_draws", C.c_int, [_ctx, C.c_uint32, u32p, u32p]),
]

_LIB = None


class AnnembedCudaError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"annembed_cuda status {status} ({STATUS.get(status, '?')}): {message}")
        self.status = status


def load(rebuild_if_stale: bool = True) -> C.CDLL:
    """Load libannembed_cuda.so (building it in-tree with nvcc if absent).  Raises if that is impossible."""
    global _LIB
    if _LIB is not None:
        return _LIB
    # ANNEMBED_CUDA_LIB (documented in README.md, "Build variants"): absolute path of another build of the SAME source
    # (annembed_b200/build.py build_library(extra_flags=..., out=...)), used by the kernel-tuning studies only
    path = os.environ.get("ANNEMBED_CUDA_LIB")
    if not path:
        path = _build.LIB
        if not os.path.exists(path) or (rebuild_if_stale and _build.needs_build()):
            path = _build.build_library()
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError if the ABI and the header drift apart
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def ptr(a: np.ndarray | None, t):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(t))
