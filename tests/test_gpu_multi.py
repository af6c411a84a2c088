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


@pytest.mark.parametrize("world,n,d,fused", [(2, 10007, 2, 0), (2, 10007, 2, 1), (2, 4096, 5, 1)])
def test_sharded_equals_single_gpu(tmp_path, world, n, d, fused):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = str(tmp_path / "multi.npz")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_worker.py"), out, str(n), str(d), str(fused)]
    subprocess.run(cmd, check=True, timeout=600, cwd=ROOT)
    multi = np.load(out)
    row_ptr, col, dst = random_graph(n, 3, 9, seed=123)
    y0 = np.random.default_rng(5).uniform(-1, 1, size=(n, d)).astype(np.float32)
    ctx = A.CudaContext(A.EmbedderParams(asked_dim=d, nb_grad_batch=4, grad_step=1.0, seed=77, dmap_init=False))
    ctx.set_graph_csr(row_ptr, col, dst)
    ctx.edge_weights(want_outputs=False)
    ctx.set_embedding(y0)
    ce0, ce1 = ctx.optimize()
    np.testing.assert_array_equal(ctx.get_embedding(), multi["y"])
    assert ctx.get_stats()["positive_samples"] == int(multi["samples"][0])
    np.testing.assert_allclose([ce0, ce1], multi["ce"], rtol=1e-12)
