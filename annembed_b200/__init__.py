"""annembed_b200 -- B200-native (sm_100a) drop-in for annembed's cross-entropy embedding optimizer.

Only the hot path lives here: `csrc/` (CUDA kernels + the C ABI of include/annembed_cuda.h) and the host-side
mirror of the reference's `Embedder` interface.  No CPU fallback exists: importing works anywhere, but any compute
call without the CUDA library and a GPU raises.
"""
from .embedparams import EmbedderParams
from .kgraph import KGraph, read_csr, write_csr
from .embedder import CudaContext, Embedder, EmbedError, KGraphProjection
from ._lib import AnnembedCudaError, load

__all__ = ["EmbedderParams", "KGraph", "read_csr", "write_csr", "CudaContext", "Embedder", "EmbedError", "KGraphProjection",
           "AnnembedCudaError", "load"]
