"""Device diffusion-map initial layout at full size (GPU box): wall time of annembed_cuda_dmap_init and of an embed()
that starts from it, on the bench workload (11M x 28 Higgs shape, k = 6)."""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
import annembed_b200 as A
import workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 11_000_000
row_ptr, col, dist = workloads.blocked_knn_graph(n, 28, 6, seed=0, device="cuda")
torch.cuda.empty_cache()
p = A.EmbedderParams(dmap_init=True, scale_rho=0.75, grad_step=1.0, nb_grad_batch=40, seed=1)
out = {"n": n}
for it in range(2):
    ctx = A.CudaContext(p)
    ctx.set_graph_csr(row_ptr, col, dist)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.dmap_init(want_output=False)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    out[f"dmap_init_ms_run{it}"] = round(1e3 * (t1 - t0), 1)
    out["sigma"] = [round(float(s), 5) for s in ctx.dmap_singular_values()[:6]]
    ctx.close()
g = A.KGraph(row_ptr, col, dist, max_nbng=6)
emb = A.Embedder(g, p)
t0 = time.perf_counter(); emb.embed(); t1 = time.perf_counter()
out["embed_with_dmap_init_ms"] = round(1e3 * (t1 - t0), 1)
out["host_phases_ms"] = {k: round(v, 1) for k, v in emb.host_timings_ms.items()}
out["cross_entropy"] = list(emb.cross_entropy)
y0 = emb.get_initial_embedding()
out["initial_layout_abs_max"] = float(np.abs(y0).max())
print(json.dumps(out))
