"""Design study (CPU only): quality of the bulk-synchronous owner-computes epoch (host build of the device code)
against the Hogwild oracle, same graph, same initial layout.  Usage: python tests/studies/semantics_study.py [n] [k] [M...]"""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np
import workloads
from oracle import oracle, quality
from tests.studies import hostsim_binding as hs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
Ms = [int(a) for a in sys.argv[3:]] or [10]

nb_batch, nbs, gs = 30, 10, 1.0
x, labels = workloads.gaussian_mixture(n, 784, seed=0)
t = time.time(); idx, dist = workloads.knn_exact(x, k); print("knn", time.time() - t)
row_ptr, col, dist = workloads.csr_from_knn(idx, dist)
scale, p = oracle.edge_weights(row_ptr, col, dist, 1.0, 1.0)
es = oracle.embedded_scales(scale)
y0 = workloads.pca_init(x, 2)
nbng = 50

def report(name, y, secs):
    q = quality.quality_stats(row_ptr, col, y, nbng)
    ce = oracle.cross_entropy(row_ptr, col, p, es, y)
    print(f"{name:28s} t={secs:6.1f}s ce={ce:.4e} nomatch={q['nb_without_match']:6d} mean_nbmatch={q['mean_nbmatch']:.3f} "
          f"pres={q['knn_preservation']:.4f} med_ratio={q['median_ratio']:.3f} mean_ratio={q['mean_ratio']:.3f}")
    return q

report("initial", y0, 0)
for seed in (1, 2):
    t = time.time(); y, done = oracle.optimize(row_ptr, col, p, es, y0, 1.0, gs, nbs, nb_batch, seed=seed); dt = time.time() - t
    report(f"oracle hogwild seed{seed}", y, dt)
for M in Ms:
    for seed in (1, 2):
        t = time.time(); y, done = hs.optimize(row_ptr, col, p, es, y0, 1.0, gs, nbs, nb_batch, M, seed); dt = time.time() - t
        report(f"bsp M={M} seed{seed} ({done/ (nbs*len(col)*nb_batch):.3f})", y, dt)
