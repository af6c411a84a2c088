"""C3 at full size: embed the 11M-node Higgs-shape graph and score the layout with the device quality estimator
(the reference cannot run its own estimator at this size: examples/higgs.rs:266-268 sub-samples to 15 %)."""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np
import annembed_b200 as A
import workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 11_000_000
row_ptr, col, dist = workloads.blocked_knn_graph(n, 28, 6, seed=0, device="cuda")
import torch; torch.cuda.empty_cache()
p = A.EmbedderParams(dmap_init=False, scale_rho=0.75, grad_step=1.0, nb_grad_batch=40, seed=1)
ctx = A.CudaContext(p)
ctx.set_graph_csr(row_ptr, col, dist)
ctx.edge_weights(want_outputs=False)
ctx.set_embedding(workloads.random_init(n, 2, seed=0))
t = time.time(); q0 = ctx.quality_estimate(100); t0 = time.time() - t
t = time.time(); ce = ctx.optimize(); t1 = time.time() - t
t = time.time(); q1 = ctx.quality_estimate(100); t2 = time.time() - t
print(json.dumps({"n": n, "optimize_s": t1, "quality_estimate_s": [t0, t2], "cross_entropy": ce,
                  "initial": {k: v for k, v in q0.items()}, "final": {k: v for k, v in q1.items()}}))
