"""T9 (CPU part): the C-ABI library builds for sm_100a, loads, exports every symbol include/annembed_cuda.h
declares, and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest
import torch

import annembed_b200 as A
from annembed_b200 import _lib

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "annembed_cuda.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(annembed_cuda_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = A.load()
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in annembed_cuda.h but not exported"
    bound = {n for n, _, _ in _lib.SYMBOLS}
    assert bound == set(names), f"ctypes binding and header differ: {bound ^ set(names)}"


def test_struct_layouts_match_header():
    # annembed_cuda_params: 2 u32, 4 f64, 6 u32 (+pad), u64, 2 u32
    assert C.sizeof(_lib.Params) == 8 + 32 + 24 + 8 + 8
    assert C.sizeof(_lib.Stats) == 22 * 8
    p = _lib.Params()
    assert A.load().annembed_cuda_default_params(C.byref(p)) == 0
    # EmbedderParams::default(), embedparams.rs:107-132
    assert (p.asked_dim, p.dmap_init, p.beta, p.b, p.scale_rho, p.grad_step) == (2, 1, 1.0, 1.0, 1.0, 2.0)
    assert (p.nb_sampling_by_edge, p.nb_grad_batch, p.grad_factor, p.hierarchy_layer, p.hubness_weighting) == (10, 20, 4, 0, 0)
    d = A.EmbedderParams()
    assert (d.asked_dim, d.dmap_init, d.beta, d.b, d.scale_rho, d.grad_step, d.nb_sampling_by_edge, d.nb_grad_batch,
            d.grad_factor, d.hierarchy_layer, d.hubness_weighting) == (2, True, 1.0, 1.0, 1.0, 2.0, 10, 20, 4, 0, False)


def test_null_arguments_are_rejected_without_a_device():
    lib = A.load()
    assert lib.annembed_cuda_default_params(None) == 1
    assert lib.annembed_cuda_create(None, None, 0) == 1
    assert lib.annembed_cuda_destroy(None) == 0
    assert lib.annembed_cuda_get_stats(None, None) == 1
    assert lib.annembed_cuda_optimize(None, None, None) == 1


def test_invalid_params_rejected():
    lib = A.load()
    p = _lib.Params()
    lib.annembed_cuda_default_params(C.byref(p))
    h = C.c_void_p()
    p.asked_dim = 0
    assert lib.annembed_cuda_create(C.byref(h), C.byref(p), 0) == 1
    p.asked_dim = 33
    assert lib.annembed_cuda_create(C.byref(h), C.byref(p), 0) == 6
    assert b"asked_dim" in lib.annembed_cuda_last_error(None)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    with pytest.raises(A.AnnembedCudaError) as e:
        A.CudaContext(A.EmbedderParams())
    assert e.value.status == 2 and "no CPU fallback" in str(e.value)
    import numpy as np
    g = A.KGraph.from_knn(np.array([[1], [2], [0], [0]]), np.ones((4, 1), np.float32))
    emb = A.Embedder(g, A.EmbedderParams(dmap_init=False))
    with pytest.raises(A.EmbedError):
        emb.embed()


def test_cpp_host_mirror_compiles_against_the_header():
    """The C++ mirror of `Embedder` (include/annembed_embedder.hpp) builds and links against the C ABI (no GPU needed)."""
    import subprocess
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp")
    subprocess.run(["make", "-C", here, "-s", "-B"], check=True)
    assert os.path.exists(os.path.join(here, "test_embedder"))
