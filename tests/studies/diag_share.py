import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np
import annembed_b200 as A
from tests.conftest import random_graph
for (d, kmax, M) in [(2, 3, 1), (2, 6, 2), (2, 6, 1), (4, 4, 1)]:
    row_ptr, col, dist = random_graph(5000, 2, kmax, seed=72)
    y0 = np.random.default_rng(2).uniform(-1, 1, size=(5000, d)).astype(np.float32)
    outs = []
    for flags in (0, 1):
        ctx = A.CudaContext(A.EmbedderParams(asked_dim=d, nb_grad_batch=3, grad_step=1.0, seed=5, flags=flags, nb_sampling_by_edge=1, mini_epochs_per_batch=M))
        ctx.set_graph_csr(row_ptr, col, dist); ctx.edge_weights(want_outputs=False); ctx.set_embedding(y0)
        ctx.optimize_batches(1, 1)
        outs.append(ctx.get_embedding())
    err = np.abs(outs[0] - outs[1]).max(axis=1)
    print((d, kmax, M), 'max', err.max(), 'q99', np.quantile(err, .99), 'median', np.median(err), 'n>1e-4', int((err > 1e-4).sum()))
