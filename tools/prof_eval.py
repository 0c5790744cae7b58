#!/usr/bin/env python3
"""One short pass of the bench workload for ncu (no timing claims are made from this run)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from noisyquantumsimulator_b200.engine import Evaluator  # noqa: E402
from noisyquantumsimulator_b200.rydberg_gates import device_prologue as dpro  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=bench.B_PER_GPU)
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
cand = bench.make_candidates(a.points, 1234)
mk, ak = dpro.split_kwargs(dict(bench.MEDIUM, **cand))
model = dpro.model_constants("jp", n_segments=bench.N_SEG, **mk)
app = dpro.pack_apparatus("jp", a.points, **ak)
ev = Evaluator(0)
d_app = torch.from_numpy(app).to("cuda:0")
for _ in range(a.iters):
    out = ev.simulate_device(model, d_app)
torch.cuda.synchronize()
print("kernel ms", ev.last_kernel_ms(), "mean F", float(out[:, 0].mean()))
