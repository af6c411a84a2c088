"""Where does the end-to-end embed() time go beyond the device loop? (GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
import annembed_b200 as A
import workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 11_000_000
row_ptr, col, dist = workloads.blocked_knn_graph(n, 28, 6, seed=0, device="cuda")
torch.cuda.empty_cache()
y0 = workloads.random_init(n, 2, seed=0)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
row_ptr = pin(row_ptr.view(np.int64)).view(np.uint64); col = pin(col.view(np.int32)).view(np.uint32); dist = pin(dist); y0 = pin(y0)
p = A.EmbedderParams(dmap_init=False, scale_rho=0.75, grad_step=1.0, nb_grad_batch=40, seed=1)
for it in range(3):
    t = [time.perf_counter()]
    def lap(): torch.cuda.synchronize(); t.append(time.perf_counter())
    ctx = A.CudaContext(p); lap()
    ctx.set_graph_csr(row_ptr, col, dist); lap()
    ctx.edge_weights(want_outputs=False); lap()
    ctx.set_embedding(y0); lap()
    ctx.optimize(want_ce=True); lap()
    y = ctx.get_embedding(); lap()
    st = ctx.get_stats()
    ctx.close(); lap()
    names = ["create", "set_graph", "edge_weights", "set_embedding", "optimize", "get_embedding", "close"]
    print(it, {k: round(1e3 * (t[i + 1] - t[i]), 1) for i, k in enumerate(names)}, "total", round(1e3 * (t[-1] - t[0]), 1),
          "dev optimize", round(st["optimize_ms"], 1), "build", round(st["build_ms"], 1))
