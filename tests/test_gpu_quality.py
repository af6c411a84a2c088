"""T7: end-to-end statistical parity of the final layout -- CUDA epoch loop vs the Hogwild oracle.

Same graph, same initial layout, same parameters (examples/mnist_digits.rs:92-100: 30 batches, grad_step 1,
10 samples/edge).  The reference is unseeded and asynchronous, so parity is distributional:
means over 20 independent runs of the quality statistics of embedder.rs:620-753 (+ kNN preservation) within 1 %
(the two ratio statistics, noisier, within 1.5 % / 2 %)."""
import numpy as np
import pytest

import annembed_b200 as A
import workloads
from oracle import oracle, quality

pytestmark = pytest.mark.gpu

# 20 runs per side: at 20 000 nodes a single run's median edge ratio moves by 1.2 % (oracle, 8 runs: 0.409 +- 0.0047), so the
# difference of two 5-run means has a standard error of 0.7 % -- the 1.5 % gate below would be a two-sigma test of the noise.
# With 20 runs it is four sigma (0.37 %); the other statistics are far inside their gates either way.
N, NBNG, RUNS = 20000, 50, 20


@pytest.fixture(scope="module")
def data():
    x, _ = workloads.gaussian_mixture(N, 784, seed=0)
    return x, workloads.pca_init(x, 2)


def mean_stats(stats):
    keys = ("nb_without_match", "mean_nbmatch", "knn_preservation", "median_ratio", "mean_ratio")
    return {k: float(np.mean([s[k] for s in stats])) for k in keys}


# (kNN, scale_rho, batches, hubness): examples/mnist_digits.rs:92-100 and examples/higgs.rs:204-211,234
@pytest.mark.parametrize("k,scale_rho,nb_batch,hub", [(10, 1.0, 30, False), (6, 0.75, 40, True)])
def test_quality_statistics_within_one_percent_of_oracle(data, k, scale_rho, nb_batch, hub):
    x, y0 = data
    idx, dist = workloads.knn_exact(x, k, device="cuda")
    row_ptr, col, dist = workloads.csr_from_knn(idx, dist)
    scale, p = oracle.edge_weights(row_ptr, col, dist, scale_rho, 1.0)
    es = oracle.embedded_scales(scale)
    neg_w = oracle.hubness_weights(row_ptr, col) if hub else None
    ref, ours = [], []
    for seed in range(RUNS):
        y, _ = oracle.optimize(row_ptr, col, p, es, y0, 1.0, 1.0, 10, nb_batch, neg_w=neg_w, seed=seed + 1)
        ref.append(quality.quality_stats(row_ptr, col, y, NBNG))
        ctx = A.CudaContext(A.EmbedderParams(nb_grad_batch=nb_batch, grad_step=1.0, scale_rho=scale_rho, seed=100 + seed,
                                             hubness_weighting=hub))
        ctx.set_graph_csr(row_ptr, col, dist)
        ctx.edge_weights(want_outputs=False)
        if hub:
            ctx.set_neg_weights(np.clip(ctx.get_hubness_counts().astype(np.float32), 1.0, float(N)))
        ctx.set_embedding(y0)
        ce0, ce1 = ctx.optimize()
        ours.append(quality.quality_stats(row_ptr, col, ctx.get_embedding(), NBNG))
        ours[-1]["ce"] = ce1
        ref[-1]["ce"] = oracle.cross_entropy(row_ptr, col, p, es, y, 1.0)
        ctx.close()
    r, o = mean_stats(ref), mean_stats(ours)
    print("oracle", r, "\ncuda  ", o)
    for k in ("mean_nbmatch", "knn_preservation"):
        assert abs(o[k] - r[k]) <= 0.01 * abs(r[k]), (k, o[k], r[k])
    # the ratio statistics of the oracle itself move by +-0.5 % between sets of 5 runs (+-0.25 % between sets of 20)
    assert abs(o["median_ratio"] - r["median_ratio"]) <= 0.015 * r["median_ratio"], (o["median_ratio"], r["median_ratio"])
    assert abs(o["mean_ratio"] - r["mean_ratio"]) <= 0.02 * r["mean_ratio"]
    # a count of rare events: within 1 % of the node count
    assert abs(o["nb_without_match"] - r["nb_without_match"]) <= 0.01 * N
    ce_r, ce_o = np.mean([s["ce"] for s in ref]), np.mean([s["ce"] for s in ours])
    assert abs(ce_o - ce_r) <= 0.05 * ce_r
