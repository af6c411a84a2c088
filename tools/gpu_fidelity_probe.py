"""GPU study: layout statistics of the CUDA epoch loop on the c3s fidelity case under different context flags /
schedules, relative to the committed oracle fixture.  usage: python tools/gpu_fidelity_probe.py [case] [runs]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import annembed_b200 as A
from tests import fidelity_cases as fc
from tests.test_gpu_fidelity import device_stats, TOL

name = sys.argv[1] if len(sys.argv) > 1 else "c3s"
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
configs = [(tuple(int(v) for v in c.split(":")) + (0,))[:3] for c in (sys.argv[3:] or ["0:0:0", "0:67:0", "0:67:1", "16:67:0"])]   # flags:M:cell_substeps
gold = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", f"fidelity_{name}.json")))
hub = name.endswith("_hub")
case = fc.make_case(name[:-4] if hub else name, n=gold["n"], device="cuda")
ref = gold["mean"]
for flags, M, S in configs:
    out = []
    for seed in range(runs):
        ctx = A.CudaContext(A.EmbedderParams(dmap_init=False, seed=1000 + seed, flags=flags, mini_epochs_per_batch=M, cell_substeps=S, hubness_weighting=hub, **case["params"]))
        ctx.set_graph_csr(case["row_ptr"], case["col"], case["dist"])
        ctx.edge_weights(want_outputs=False)
        if hub:
            ctx.set_neg_weights(np.clip(ctx.get_hubness_counts().astype(np.float32), 1.0, float(case["n"])))
        ctx.set_embedding(case["y0"])
        ce0, ce1 = ctx.optimize()
        s = device_stats(ctx, case, ctx.get_embedding()); s["ce"] = ce1
        out.append(s)
        st = ctx.get_stats(); ms = st["optimize_ms"]
        ctx.close()
    mean = {k: float(np.mean([r[k] for r in out])) for k in TOL}
    print(f"{name} flags={flags} M={M} S={S}->{st['cell_substeps']} cells={st['n_cells']} cross={st['cross_cell_edges']} optimize_ms={ms:.0f}", {k: round(mean[k] / ref[k] - 1, 4) for k in TOL}, flush=True)
