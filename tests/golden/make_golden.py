"""Generates tests/golden/hotpath_small.npz from the CPU oracle (oracle/annembed_oracle.c).

The Rust reference cannot run here and holds no golden vectors for this path (F5), so these
vectors pin the ORACLE's outputs; they are what the CUDA path must reproduce.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from tests.conftest import random_graph  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hotpath_small.npz")


def main():
    row_ptr, col, dist = random_graph(200, 2, 11, seed=2024, zero_frac=0.06, dup_rows=4)
    n = 200
    out = dict(row_ptr=row_ptr, col=col, dist=dist)
    for tag, (rho, beta) in {"a": (1.0, 1.0), "b": (0.75, 2.0), "c": (1.5, 0.5)}.items():
        s, p = oracle.edge_weights(row_ptr, col, dist, rho, beta)
        out[f"w_{tag}_params"] = np.array([rho, beta])
        out[f"w_{tag}_scale"] = s
        out[f"w_{tag}_p"] = p
        out[f"w_{tag}_perplexity"] = oracle.perplexity(row_ptr, p)
    scale, p = out["w_a_scale"], out["w_a_p"]
    es = oracle.embedded_scales(scale)
    out["emb_scale"] = es
    rng = np.random.Generator(np.random.PCG64(11))
    for d in (2, 3, 15):
        y0 = rng.uniform(-0.5, 0.5, size=(n, d)).astype(np.float32)
        y0[5] = y0[9]                      # coincident points: zero-distance branches
        edges = rng.integers(0, len(col), size=64).astype(np.uint64)
        negs = rng.integers(0, n, size=(64, 5)).astype(np.uint32)
        negs[3, :2] = 5; edges[3] = int(row_ptr[9])      # node 9's first edge with negatives coincident with it
        out[f"y0_d{d}"] = y0
        out[f"edges_d{d}"] = edges
        out[f"negs_d{d}"] = negs
        for b in (1.0, 0.5):
            for gs in (1.0, 0.05):
                out[f"step_d{d}_b{b}_g{gs}"] = oracle.step_fixed(row_ptr, col, p, es, y0, b, gs, edges, negs)
            out[f"ce_d{d}_b{b}"] = np.array(oracle.cross_entropy(row_ptr, col, p, es, y0, b))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
