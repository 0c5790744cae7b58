#!/usr/bin/env python3
"""5-level model (config 5): one batch of LP points with distinct second-leg powers, evaluated from
device-resident rows.  Prints kernel time and achieved FP64 rate; also the pass ncu captures."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from noisyquantumsimulator_b200.engine import Evaluator  # noqa: E402
from noisyquantumsimulator_b200.rydberg_gates import prologue  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=65536)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--species", default="Rb87")
ap.add_argument("--noiseless", action="store_true")
a = ap.parse_args()
B = a.points
p2 = 10 ** np.random.default_rng(1234).uniform(np.log10(0.05), np.log10(2.0), B)
kw = dict(bench.MEDIUM, species=a.species, laser_2_power=p2, hilbert_space_dim=5, include_noise=not a.noiseless)
plan = prologue.build_plan("lp", B=B, warn=False, **kw)
ev = Evaluator(0)
dev = torch.device("cuda:0")
d_pts, d_seg = torch.from_numpy(plan.points).to(dev), torch.from_numpy(plan.segs).to(dev)
d_ns = torch.from_numpy(plan.n_seg).to(dev)
peak, _ = ev.fp64_peak()
ms = []
for _ in range(a.iters):
    out, _ = ev.evaluate_device(5, not a.noiseless, d_pts, d_seg, d_ns)
    torch.cuda.synchronize()
    ms.append(ev.last_kernel_ms())
o = out.cpu().numpy()
k = min(ms[1:] or ms)
fl = float(o[:, 15].sum())
print(f"dim5 {a.species} B={B} kernel {k:.3f} ms  {B / k * 1e3 / 1e6:.3f} M evals/s  {fl / k * 1e-9:.2f} TFLOP/s "
      f"({fl / k * 1e-9 / peak * 100:.1f}% of DFMA peak {peak:.1f})  mean F {o[:, 0].mean():.6f}  "
      f"flops/point {fl / B / 1e6:.2f} M  steps/point {o[:, 14].mean():.0f}")
