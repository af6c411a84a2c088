"""The full C3 workload (11M nodes, k=6, random init, 40 batches) embedded once by the CUDA path and once by the oracle's
Hogwild loop on the box's host cores, same graph, same initial layout; both layouts scored by the same device quality
estimator (annembed_cuda_quality_estimate, nbng 100) and the same K5.  Writes gpurun_out/full_c3_vs_oracle.json.
usage: python tools/gpu_full_c3_vs_oracle.py [nodes] [oracle_batches]   (about 15 minutes of host CPU at 11M nodes)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import annembed_b200 as A
import workloads
from oracle import oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 11_000_000
row_ptr, col, dist = workloads.blocked_knn_graph(n, 28, 6, seed=0, device="cuda")
y0 = workloads.random_init(n, 2, seed=0)
prm = dict(asked_dim=2, nb_grad_batch=40, scale_rho=0.75, grad_step=1.0)
KEYS = ("nb_without_match", "mean_nbmatch", "knn_preservation", "mean_ratio")


def score(ctx, y):
    ctx.set_embedding(y)
    q = ctx.quality_estimate(100)
    s = {k: float(q[k]) for k in KEYS}
    s["median_ratio"] = float(q["ratio_quantiles"][2])
    s["ce"] = float(ctx.cross_entropy())
    return s


ctx = A.CudaContext(A.EmbedderParams(dmap_init=False, seed=2024, **prm))
ctx.set_graph_csr(row_ptr, col, dist)
scale, p = ctx.edge_weights()
es = ctx.get_embedded_scales()
ctx.set_embedding(y0)
t = time.time(); ce0, ce1 = ctx.optimize(); t_gpu = time.time() - t
y_gpu = ctx.get_embedding()
st = ctx.get_stats()
out = {"nodes": n, "params": prm, "cuda": score(ctx, y_gpu), "cuda_seconds": t_gpu, "cuda_optimize_ms": st["optimize_ms"], "ce_initial": ce0}
print("cuda", out["cuda"], flush=True)
t = time.time()
y_or, done = oracle.optimize(row_ptr, col, p, es, y0, 1.0, prm["grad_step"], 10, prm["nb_grad_batch"], seed=7)
out["oracle_seconds"] = time.time() - t
out["oracle_threads"] = oracle.num_threads()
out["oracle"] = score(ctx, y_or)
out["relative_difference"] = {k: out["cuda"][k] / out["oracle"][k] - 1 for k in out["oracle"]}
print("oracle", out["oracle"], out["oracle_seconds"], flush=True)
print("relative difference", out["relative_difference"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/full_c3_vs_oracle.json", "w"), indent=1)
