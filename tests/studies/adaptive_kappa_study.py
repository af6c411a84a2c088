"""CPU study (host build of the device code): does a coarser mini-epoch (larger kappa) in the late, small-step batches
keep the layout statistics within 1 % of the oracle?  Usage: python tests/studies/adaptive_kappa_study.py"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np
import workloads
from oracle import oracle, quality
from tests.studies import hostsim_binding as hs

n, k, nb, nbs, gs = 20000, 10, 30, 10, 1.0
x, _ = workloads.gaussian_mixture(n, 784, seed=0)
idx, dist = workloads.knn_exact(x, k)
row_ptr, col, dist = workloads.csr_from_knn(idx, dist)
scale, p = oracle.edge_weights(row_ptr, col, dist)
es = oracle.embedded_scales(scale)
y0 = workloads.pca_init(x, 2)
keys = ("nb_without_match", "mean_nbmatch", "knn_preservation", "median_ratio", "mean_ratio")

def summarize(name, ys):
    st = [quality.quality_stats(row_ptr, col, y, 50) for y in ys]
    print(f"{name:34s} " + " ".join(f"{kk}={np.mean([s[kk] for s in st]):.4f}" for kk in keys), flush=True)


schedules = {
    "M=34 throughout": [(1, 30, 34)],
    "M=34 (1-15), 17 (16-30)": [(1, 15, 34), (16, 15, 17)],
    "M=34 (1-10), 17 (11-20), 9 (21-30)": [(1, 10, 34), (11, 10, 17), (21, 10, 9)],
    "M=17 throughout": [(1, 30, 17)],
    "M=17 (1-15), 34 (16-30)": [(1, 15, 17), (16, 15, 34)],
    "M=9 (1-10), 17 (11-20), 34 (21-30)": [(1, 10, 9), (11, 10, 17), (21, 10, 34)],
    "M=17 (1-20), 50 (21-30)": [(1, 20, 17), (21, 10, 50)],
    "M=10 (1-15), 34 (16-30)": [(1, 15, 10), (16, 15, 34)],
}
schedules = {k: v for k, v in schedules.items() if "(1-1" in k or "(1-2" in k}
for name, sched in schedules.items():
    ys = []
    for seed in range(4):
        y = y0
        for (first, count, M) in sched:
            y, _ = hs.optimize(row_ptr, col, p, es, y, 1.0, gs, nbs, nb, M, seed + 10, None, first, count)
        ys.append(y)
    summarize(name, ys)
