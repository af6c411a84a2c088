"""CPU study (host build of the device code): how coarse can the EARLY batches of the graded mini-epoch schedule be
while the final layout statistics stay within 1 % of the serial oracle's?  Two parameter sets, as the GPU gate
(tests/test_gpu_quality.py).  Usage: python tests/studies/schedule_study.py [mnist|higgs] [seeds]"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np
import workloads
from oracle import oracle, quality
from tests.studies import hostsim_binding as hs
from tests.test_host import _vose_alias

which = sys.argv[1] if len(sys.argv) > 1 else "mnist"
SEEDS = int(sys.argv[2]) if len(sys.argv) > 2 else 6
if which == "mnist":
    n, k, nb, rho, hub = 20000, 10, 30, 1.0, False
else:
    n, k, nb, rho, hub = 20000, 6, 40, 0.75, True
nbs, gs = 10, 1.0
x, _ = workloads.gaussian_mixture(n, 784 if which == "mnist" else 28, seed=0)
idx, dist = workloads.knn_exact(x, k)
row_ptr, col, dist = workloads.csr_from_knn(idx, dist)
scale, p = oracle.edge_weights(row_ptr, col, dist, rho, 1.0)
es = oracle.embedded_scales(scale)
y0 = workloads.pca_init(x, 2)
neg_w = oracle.hubness_weights(row_ptr, col) if hub else None
tab = _vose_alias(neg_w) if hub else None
keys = ("nb_without_match", "mean_nbmatch", "knn_preservation", "median_ratio", "mean_ratio")
base = {}


def summarize(name, ys, secs):
    st = [quality.quality_stats(row_ptr, col, y, 50) for y in ys]
    m = {kk: float(np.mean([s[kk] for s in st])) for kk in keys}
    sd = {kk: float(np.std([s[kk] for s in st]) / np.sqrt(len(st))) for kk in keys}
    if not base:
        base.update(m)
    print(f"{name:30s} t={secs:5.1f}s " + " ".join(f"{kk}={m[kk]:.4f}({100 * (m[kk] / base[kk] - 1):+.2f}%,se{100 * sd[kk] / base[kk]:.2f})" for kk in keys[1:]), flush=True)


t = time.time()
ys = [oracle.optimize(row_ptr, col, p, es, y0, 1.0, gs, nbs, nb, neg_w=neg_w, seed=s + 1)[0] for s in range(SEEDS)]
summarize("oracle hogwild", ys, time.time() - t)
third = nb // 3
def graded(a, b, c):
    # same thirds as mini_epochs_of_batch (annembed_cuda.cu): 3*iter <= nb, 3*iter <= 2*nb, rest
    i1 = nb // 3; i2 = (2 * nb) // 3
    return [(1, i1, a), (i1 + 1, i2 - i1, b), (i2 + 1, nb - i2, c)]
schedules = {
    "34 throughout": [(1, nb, 34)],
    "9 / 17 / 34 (shipped)": graded(9, 17, 34),
    "5 / 9 / 34": graded(5, 9, 34),
    "5 / 17 / 34": graded(5, 17, 34),
    "3 / 9 / 34": graded(3, 9, 34),
    "9 / 9 / 34": graded(9, 9, 34),
}
for name, sched in schedules.items():
    t = time.time(); ys = []
    for seed in range(SEEDS):
        y = y0
        for (first, count, M) in sched:
            y, _ = hs.optimize(row_ptr, col, p, es, y, 1.0, gs, nbs, nb, M, seed + 10, tab, first, count)
        ys.append(y)
    summarize(name, ys, time.time() - t)
