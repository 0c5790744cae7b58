#!/usr/bin/env python3
"""One short pass of the bench workload for ncu (no timing claims are made from this run)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from noisyquantumsimulator_b200.engine import Evaluator  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=65536)
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
pts, segs, _ = bench.make_candidates(a.points, 1234)
ev = Evaluator(0)
dev = torch.device("cuda:0")
d_pts, d_segs = torch.from_numpy(pts).to(dev), torch.from_numpy(segs).to(dev)
for _ in range(a.iters):
    out, _ = ev.evaluate_device(3, True, d_pts, d_segs, None)
torch.cuda.synchronize()
print("kernel ms", ev.last_kernel_ms(), "mean F", float(out[:, 0].mean()))
