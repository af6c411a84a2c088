"""Design study (CPU only): quality statistics of the bulk-synchronous epoch loop (host build of the device code,
graded schedule as in annembed_cuda.cu) against the Hogwild oracle on the BASELINE.json configs.

Usage: python tests/studies/fidelity_configs.py <c1|c2|c3s|c4s> [n] [runs] [--hub] [--oracle-only|--bsp-only]
Prints one JSON line per run and a summary; used to choose the schedule, the tests proper are tests/test_gpu_fidelity.py."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np

import workloads
from oracle import oracle, quality
from tests.studies import hostsim_binding as hs
from tests import fidelity_cases as fc
from tests.test_host import _vose_alias


def graded_bsp(row_ptr, col, p, es, y0, nbs, nb_batch, seed, neg_alias=None, spe=0.3, div=(4, 2, 1)):
    base = int(np.ceil(nbs / spe))
    y = y0
    for it in range(1, nb_batch + 1):
        dv = div[0] if 3 * it <= nb_batch else (div[1] if 3 * it <= 2 * nb_batch else div[2])
        M = max(1, (base + dv - 1) // dv)
        y, _ = hs.optimize(row_ptr, col, p, es, y, 1.0, 1.0, nbs, nb_batch, M, seed, neg_alias=neg_alias, first_batch=it, n_batches=1)
    return y


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    flags = [a for a in sys.argv[1:] if a.startswith("--")]
    name = args[0]
    n = int(args[1]) if len(args) > 1 else None
    runs = int(args[2]) if len(args) > 2 else 5
    hub = "--hub" in flags
    case = fc.make_case(name, n=n, device="cpu")
    row_ptr, col, dist, y0, prm = case["row_ptr"], case["col"], case["dist"], case["y0"], case["params"]
    scale, p = oracle.edge_weights(row_ptr, col, dist, prm["scale_rho"], 1.0)
    es = oracle.embedded_scales(scale)
    neg_w = oracle.hubness_weights(row_ptr, col) if hub else None
    alias = _vose_alias(neg_w) if hub else None
    out = {"oracle": [], "bsp": []}
    for seed in range(runs):
        if "--bsp-only" not in flags:
            t = time.time()
            y, _ = oracle.optimize(row_ptr, col, p, es, y0, 1.0, prm["grad_step"], 10, prm["nb_grad_batch"], neg_w=neg_w, seed=seed + 1)
            q = fc.summary(quality.quality_stats(row_ptr, col, y, case["nbng"]))
            q["ce"] = oracle.cross_entropy(row_ptr, col, p, es, y, 1.0); q["secs"] = time.time() - t
            out["oracle"].append(q); print("oracle", seed, json.dumps(q), flush=True)
        if "--oracle-only" not in flags:
            t = time.time()
            y = graded_bsp(row_ptr, col, p, es, y0, 10, prm["nb_grad_batch"], 100 + seed, neg_alias=alias)
            q = fc.summary(quality.quality_stats(row_ptr, col, y, case["nbng"]))
            q["ce"] = oracle.cross_entropy(row_ptr, col, p, es, y, 1.0); q["secs"] = time.time() - t
            out["bsp"].append(q); print("bsp   ", seed, json.dumps(q), flush=True)
    for k, v in out.items():
        if v:
            print(k, "mean", json.dumps({kk: float(np.mean([s[kk] for s in v])) for kk in v[0]}))


if __name__ == "__main__":
    main()
