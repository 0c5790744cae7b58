#!/usr/bin/env python3
"""One Monte-Carlo sweep (16 x 16 x 16 grid, 64 shots per point) for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from noisyquantumsimulator_b200 import rydberg_gates as rg  # noqa: E402
from noisyquantumsimulator_b200.engine import get_evaluator  # noqa: E402

ev = get_evaluator(0)
p2, dg, ot = np.meshgrid(np.geomspace(0.05, 2.0, 16), np.linspace(0.20, 0.50, 16), np.linspace(3.5, 5.5, 16), indexing="ij")
kw = dict(bench.MEDIUM, laser_2_power=p2.ravel(), delta_over_omega=dg.ravel(), omega_tau=ot.ravel(), temperature=10e-6)
for _ in range(2):
    r = rg.simulate_CZ_gate_batch("lp", B=p2.size, n_shots=64, seed=1234, warn=False, copy=False, **kw)
print("kernel span ms", ev.last_kernel_ms(), "coef ms", ev.last_coef_ms(), "cluster stats", ev.cluster_stats(),
      "mean F", float(r.avg_fidelity.mean()))
