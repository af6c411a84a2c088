"""Generates tests/golden/fidelity_<case>.json: quality statistics of the ORACLE's layouts (oracle/annembed_oracle.c, the
restatement of embedder.rs:794-1315, Hogwild on all host cores) on the BASELINE.json configs, 5 seeds each.

The GPU tests (tests/test_gpu_fidelity.py) rebuild the same seeded case on the test box, run the CUDA path 5 times and
compare the means of the same statistics at the north-star tolerance (1 %).  Running the oracle here instead of on the
GPU box keeps the GPU suite short (a 1M-node, 40-batch oracle embed is 2-3 minutes of CPU per seed).

Usage: python tests/golden/make_fidelity_golden.py <case>[:hub] ...      cases: tests/fidelity_cases.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np

from oracle import oracle, quality
from tests import fidelity_cases as fc

HERE = os.path.dirname(os.path.abspath(__file__))
RUNS = 5
SIZES = {"c1": 70000, "c2": 70000, "c3s": 1_000_000, "c4s": 200_000}


def main():
    for spec in sys.argv[1:]:
        name, _, opt = spec.partition(":")
        hub = opt == "hub"
        case = fc.make_case(name, n=SIZES[name], device="cpu")
        row_ptr, col, dist, y0, prm = case["row_ptr"], case["col"], case["dist"], case["y0"], case["params"]
        scale, p = oracle.edge_weights(row_ptr, col, dist, prm["scale_rho"], 1.0)
        es = oracle.embedded_scales(scale)
        neg_w = oracle.hubness_weights(row_ptr, col) if hub else None
        runs = []
        for seed in range(RUNS):
            t = time.time()
            y, _ = oracle.optimize(row_ptr, col, p, es, y0, 1.0, prm["grad_step"], 10, prm["nb_grad_batch"], neg_w=neg_w, seed=seed + 1)
            q = fc.summary(fc.quality_stats(row_ptr, col, y, case["nbng"]))
            q["ce"] = float(oracle.cross_entropy(row_ptr, col, p, es, y, 1.0))
            q["seconds"] = round(time.time() - t, 1)
            runs.append(q)
            print(spec, seed, json.dumps(q), flush=True)
        out = {
            "case": name, "hubness": hub, "n": case["n"], "k": case["k"], "nbng": case["nbng"], "params": prm, "runs": runs,
            "mean": {k: float(np.mean([r[k] for r in runs])) for k in runs[0] if k != "seconds"},
            "generator": "tests/golden/make_fidelity_golden.py (oracle/annembed_oracle.c, %d OpenMP threads)" % oracle.num_threads(),
            # the first rows of the graph, so that the test can verify that it rebuilt the same case
            "col_head": [int(c) for c in col[: 64 * case["k"]]],
            "sum_dist": float(np.sum(dist, dtype=np.float64)),
        }
        path = os.path.join(HERE, f"fidelity_{name}{'_hub' if hub else ''}.json")
        with open(path, "w") as f:
            json.dump(out, f, indent=1)
        print("wrote", path, json.dumps(out["mean"]))


if __name__ == "__main__":
    main()
