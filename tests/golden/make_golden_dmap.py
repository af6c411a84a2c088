"""Generates tests/golden/dmap_small.npz from the CPU restatement of the diffusion-map initial layout (oracle/dmap.py)
on the graph of hotpath_small.npz (ragged rows, zero distances, duplicated rows).  Pins the ORACLE's outputs (the Rust
reference cannot run here and holds no vectors for this path).  Run:  python tests/golden/make_golden_dmap.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dmap  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    G = np.load(os.path.join(HERE, "hotpath_small.npz"))
    row_ptr, col, dist = G["row_ptr"], G["col"], G["dist"]
    n = len(row_ptr) - 1
    vd, v, sw, normed = dmap.sym_kernel(row_ptr, col, dist, 12)
    omega = np.random.Generator(np.random.PCG64(5)).standard_normal((n, 20)).astype(np.float32)
    out = dict(diag=vd, val=v, sw=sw, normed=normed, omega=omega)
    for d in (2, 3):
        y, lam, U = dmap.dmap_layout_randomized(row_ptr, col, dist, asked_dim=d, omega=omega)
        out[f"layout_d{d}"] = y
        out["sigma"] = lam
    np.savez_compressed(os.path.join(HERE, "dmap_small.npz"), **out)
    print("wrote dmap_small.npz", os.path.getsize(os.path.join(HERE, "dmap_small.npz")), "bytes")


if __name__ == "__main__":
    main()
