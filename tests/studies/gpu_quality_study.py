"""GPU study: layout-quality statistics of the CUDA epoch loop vs the Hogwild oracle for several mini-epoch counts.
Usage (GPU box): python tests/studies/gpu_quality_study.py [n] [k] [M ...]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np
import annembed_b200 as A
import workloads
from oracle import oracle, quality

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
Ms = [int(a) for a in sys.argv[3:]] or [0]
RUNS, NBNG = 4, 50
FLAGS = int(os.environ.get('ANNEMBED_FLAGS', '0'))
RHO = float(os.environ.get('SCALE_RHO', '1.0')); NB = int(os.environ.get('NB_BATCH', '30')); HUB = int(os.environ.get('HUB', '0'))
x, _ = workloads.gaussian_mixture(n, 784, seed=0)
idx, dist = workloads.knn_exact(x, k, device="cuda")
row_ptr, col, dist = workloads.csr_from_knn(idx, dist)
scale, p = oracle.edge_weights(row_ptr, col, dist, RHO, 1.0)
neg_w = oracle.hubness_weights(row_ptr, col) if HUB else None
es = oracle.embedded_scales(scale)
y0 = workloads.pca_init(x, 2)
keys = ("nb_without_match", "mean_nbmatch", "knn_preservation", "median_ratio", "mean_ratio")

def summarize(name, ys, secs):
    st = [quality.quality_stats(row_ptr, col, y, NBNG) for y in ys]
    ce = np.mean([oracle.cross_entropy(row_ptr, col, p, es, y) for y in ys])
    print(f"{name:22s} t={secs:6.2f}s ce={ce:.4e} " + " ".join(f"{kk}={np.mean([s[kk] for s in st]):.4f}" for kk in keys), flush=True)

t = time.time(); ys = [oracle.optimize(row_ptr, col, p, es, y0, 1.0, 1.0, 10, NB, neg_w=neg_w, seed=s + 1)[0] for s in range(RUNS)]
summarize("oracle hogwild", ys, (time.time() - t) / RUNS)
for M in Ms:
    ys = []; t = time.time()
    for s in range(RUNS):
        ctx = A.CudaContext(A.EmbedderParams(nb_grad_batch=NB, grad_step=1.0, scale_rho=RHO, seed=100 + s, mini_epochs_per_batch=M, flags=FLAGS, hubness_weighting=bool(HUB)))
        ctx.set_graph_csr(row_ptr, col, dist); ctx.edge_weights(want_outputs=False)
        if HUB: ctx.set_neg_weights(neg_w)
        ctx.set_embedding(y0)
        ctx.optimize(want_ce=False); ys.append(ctx.get_embedding()); Meff = ctx.get_stats()["mini_epochs_per_batch"]; ctx.close()
    summarize(f"cuda M={M if M else 'graded(' + str(Meff) + ')'} flags={FLAGS}", ys, (time.time() - t) / RUNS)
