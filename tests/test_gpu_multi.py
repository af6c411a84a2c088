"""T8: node-sharded runs over R GPUs give bit-identical layouts to R = 1 (owner-computes is
deterministic and the draws are keyed by (seed, node, firing, epoch), not by rank), with either exchange:
fused = 0 NCCL all-gather, fused = 1 peer-memory row stores from the in-edge kernel + barrier."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

import annembed_b200 as A
from tests.conftest import random_graph

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


BULK = 32   # ANNEMBED_FLAG_BULK_SYNCHRONOUS


def _run_workers(tmp_path, world, n, d, fused, flags):
    out = str(tmp_path / "multi.npz")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_worker.py"), out, str(n), str(d), str(fused), str(flags)]
    subprocess.run(cmd, check=True, timeout=600, cwd=ROOT)
    return np.load(out)


def _single(n, d, flags):
    row_ptr, col, dst = random_graph(n, 3, 9, seed=123)
    y0 = np.random.default_rng(5).uniform(-1, 1, size=(n, d)).astype(np.float32)
    ctx = A.CudaContext(A.EmbedderParams(asked_dim=d, nb_grad_batch=4, grad_step=1.0, seed=77, dmap_init=False, flags=flags))
    ctx.set_graph_csr(row_ptr, col, dst)
    ctx.edge_weights(want_outputs=False)
    ctx.set_embedding(y0)
    ce0, ce1 = ctx.optimize()
    return ctx.get_embedding(), ctx.get_stats(), (ce0, ce1), len(col)


# fused = 0: no peer memory -> the bulk-synchronous form with the NCCL all-gather, whatever the flags
@pytest.mark.parametrize("world,n,d,fused,flags", [(2, 10007, 2, 0, 0), (2, 10007, 2, 1, BULK), (2, 4096, 5, 1, BULK)])
def test_bulk_synchronous_sharded_equals_single_gpu(tmp_path, world, n, d, fused, flags):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    multi = _run_workers(tmp_path, world, n, d, fused, flags)
    y, st, ce, _ = _single(n, d, BULK)
    np.testing.assert_array_equal(y, multi["y"])
    assert st["positive_samples"] == int(multi["samples"][0])
    np.testing.assert_allclose(ce, multi["ce"], rtol=1e-12)


@pytest.mark.parametrize("world,n,d", [(2, 20000, 2), (2, 6000, 5)])
def test_asynchronous_sharded_is_the_same_optimisation(tmp_path, world, n, d):
    """The default on several ranks (with peer memory): every rank sweeps its own part asynchronously, moves of nodes owned
    elsewhere are reduced into the owner's replica over NVLink, the owners' rows are exchanged every few launches.
    Another realisation of the single-GPU optimisation: same sample count to a fraction of a per cent, close cross entropy."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    multi = _run_workers(tmp_path, world, n, d, 1, 0)
    y, st, ce, E = _single(n, d, 0)
    assert np.isfinite(multi["y"]).all() and multi["y"].shape == y.shape
    assert abs(float(multi["samples"][0]) / st["positive_samples"] - 1) < 0.02
    assert abs(float(multi["samples"][0]) / (4 * 10 * E) - 1) < 0.02
    assert abs(multi["ce"][0] / ce[0] - 1) < 1e-12                     # the same initial layout, the same K5 (fp64 partial sums per rank)
    assert abs(multi["ce"][1] / ce[1] - 1) < 0.03 and multi["ce"][1] < multi["ce"][0]
    assert 0 < int(multi["exchanges"]) <= int(multi["launches"]) and int(multi["cross_rank_edges"]) > 0
