"""GPU study: piecewise-constant mini-epoch schedules on a fidelity case.  One context per segment (the layout is carried
from one to the next through the C ABI), same gradient-step schedule as one 40-batch run.
usage: python tools/gpu_schedule_probe.py <case> <runs> <spec> ...   spec = M1xB1,M2xB2,...  (sum of B = nb_grad_batch)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import annembed_b200 as A
from tests import fidelity_cases as fc
from tests.test_gpu_fidelity import device_stats, TOL

name, runs, specs = sys.argv[1], int(sys.argv[2]), sys.argv[3:]
hub = name.endswith("_hub")
base = name[:-4] if hub else name
gold = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", f"fidelity_{name}.json")))
case = fc.make_case(base, n=gold["n"], device="cuda")
ref = gold["mean"]
for spec in specs:
    segs = [tuple(int(v) for v in s.split("x")) for s in spec.split(",")]
    assert sum(b for _, b in segs) == case["params"]["nb_grad_batch"], spec
    out, ms = [], 0.0
    for seed in range(runs):
        y = case["y0"]
        first = 1
        tot = 0.0
        for M, B in segs:
            ctx = A.CudaContext(A.EmbedderParams(dmap_init=False, seed=1000 + seed, mini_epochs_per_batch=M, hubness_weighting=hub, **case["params"]))
            ctx.set_graph_csr(case["row_ptr"], case["col"], case["dist"])
            ctx.edge_weights(want_outputs=False)
            if hub:
                ctx.set_neg_weights(np.clip(ctx.get_hubness_counts().astype(np.float32), 1.0, float(case["n"])))
            ctx.set_embedding(y)
            ctx.optimize_batches(first, B)
            tot += ctx.get_stats()["optimize_ms"]
            first += B
            y = ctx.get_embedding()
            last = ctx
        s = device_stats(last, case, y); s["ce"] = last.cross_entropy()
        out.append(s); ms = tot
    mean = {k: float(np.mean([r[k] for r in out])) for k in TOL}
    print(f"{name} schedule {spec} optimize_ms={ms:.0f}", {k: round(mean[k] / ref[k] - 1, 4) for k in TOL}, flush=True)
