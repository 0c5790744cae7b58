#!/usr/bin/env python3
"""One smooth-JP batch (65,536 candidates inside the optimiser's bounds) for ncu: the repeated-segment path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from noisyquantumsimulator_b200.engine import get_evaluator  # noqa: E402
from noisyquantumsimulator_b200.rydberg_gates import batched_objective as bo  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ev = get_evaluator(0)
bounds, x0, _ = bo.bounds_and_x0("sjp", 5)
lo, hi = np.array(bounds).T
X = lo[:, None] + (hi - lo)[:, None] * np.random.default_rng(43).uniform(size=(len(bounds), n))
f = bo.make_vectorized_objective("sjp", include_noise=True, n_segments=5)
for _ in range(2):
    c = f(X)
print("kernel span ms", ev.last_kernel_ms(), "repeated-segment points", ev.repeat_stats(), "best cost", float(np.min(c)))
