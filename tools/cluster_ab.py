#!/usr/bin/env python3
"""A/B of the clustered-spectrum path (csrc/rydcz_cluster.cuh) against the plain Chebyshev path on the
bench workload: same batch, cluster mode 0 / 1 / 2, device time per pass, kernel and coefficient-kernel
times, how many points took the path, largest difference of the result rows."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from noisyquantumsimulator_b200.engine import Evaluator  # noqa: E402
from noisyquantumsimulator_b200.rydberg_gates import device_prologue as dpro  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=bench.B_PER_GPU)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--modes", default="0,1")
ap.add_argument("--tables", default="1", help="shared coefficient table modes to run with every cluster mode > 0")
ap.add_argument("--noiseless", action="store_true")
a = ap.parse_args()
cand = bench.make_candidates(a.points, 1234)
medium = dict(bench.MEDIUM)
if a.noiseless:
    medium["include_noise"] = False
mk, ak = dpro.split_kwargs(dict(medium, **cand))
model = dpro.model_constants("jp", n_segments=bench.N_SEG, **mk)
app = dpro.pack_apparatus("jp", a.points, **ak)
ev = Evaluator(0)
d_app = torch.from_numpy(app).to("cuda:0")
ref = None
runs = []
for mode in [int(m) for m in a.modes.split(",")]:
    runs += [(mode, 1)] if mode == 0 else [(mode, int(t)) for t in a.tables.split(",")]
for mode, table in runs:
    ev.set_cluster(mode)
    ev.set_table(table)
    for _ in range(2):
        out = ev.simulate_device(model, d_app)
    torch.cuda.synchronize()
    ev.cluster_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        out = ev.simulate_device(model, d_app)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    planned, back = ev.cluster_stats()
    gen, shared = ev.table_stats()
    o = out.cpu().numpy()
    if ref is None:
        ref = o
    line = {"mode": mode, "table": table, "table_entries_per_pass": gen / a.iters, "table_slots_per_pass": shared / a.iters,
            "points": a.points, "ms_per_pass": ms, "evals_per_s": a.points / ms * 1e3,
            "kernel_ms": ev.last_kernel_ms(), "coef_ms": ev.last_coef_ms(),
            "planned_per_pass": planned / a.iters, "handed_back": back,
            "max_diff_vs_first_mode": float(np.nanmax(np.abs(o[:, :14] - ref[:, :14]))),
            "mean_F": float(o[:, 0].mean()), "flops": float(o[:, 15].sum()), "steps": float(o[:, 14].sum())}
    print(json.dumps(line), flush=True)
