"""Small end-to-end script for `compute-sanitizer --tool memcheck|racecheck python tests/sanitizer_script.py` (T9)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import annembed_b200 as A
from tests.conftest import random_graph

# flags: 4 = no relabelling (debug_draws reports by edge of the caller's graph), 8 = replayed in-edge decisions
for (n, kmin, kmax, d, hub, flags, mini) in [(1003, 2, 7, 2, False, 4, 0), (1777, 3, 14, 3, True, 0, 0), (500, 17, 20, 2, False, 4, 0), (1900, 6, 6, 15, False, 0, 0),
                                           (1290, 6, 6, 2, False, 8, 3), (1650, 8, 8, 4, True, 4, 2), (2100, 4, 10, 2, False, 0, 5)]:   # constant row length, many firings per node
    row_ptr, col, dist = random_graph(n, kmin, kmax, seed=n)
    ctx = A.CudaContext(A.EmbedderParams(asked_dim=d, nb_grad_batch=2, grad_step=1.0, hubness_weighting=hub, flags=flags,
                                           mini_epochs_per_batch=mini))
    ctx.set_graph_csr(row_ptr, col, dist)
    ctx.edge_weights()
    if hub:
        c = ctx.get_hubness_counts()
        ctx.set_neg_weights(np.clip(c.astype(np.float32), 1, n))
    ctx.set_embedding(np.random.default_rng(0).uniform(-1, 1, (n, d)).astype(np.float32))
    ctx.step_fixed(np.arange(5, dtype=np.uint64), np.random.default_rng(1).integers(0, n, (5, 5)).astype(np.uint32), 0.5)
    ce = ctx.optimize()
    if flags & 4:
        ctx.debug_draws(1)
    y = ctx.get_embedding()
    assert np.isfinite(y).all()
    ctx.close()
print("sanitizer script ok")
