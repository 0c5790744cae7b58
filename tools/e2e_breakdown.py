"""Where the end-to-end time of simulate_CZ_gate_batch goes (host arrays in -> host arrays out)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from noisyquantumsimulator_b200 import rydberg_gates as rg  # noqa: E402
from noisyquantumsimulator_b200.engine import get_evaluator  # noqa: E402
from noisyquantumsimulator_b200.rydberg_gates import device_prologue as dpro  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
cand = bench.make_candidates(B, 1234)
kw = dict(bench.MEDIUM, **cand)
ev = get_evaluator(0)
for _ in range(3):
    rg.simulate_CZ_gate_batch("jp", B=B, prologue="device", copy=False, **kw)
T = {"total": [], "columns+model": [], "c_call": [], "kernel_span": []}
for _ in range(8):
    t0 = time.perf_counter()
    r = rg.simulate_CZ_gate_batch("jp", B=B, prologue="device", copy=False, **kw)
    t1 = time.perf_counter()
    mk, ak = dpro.split_kwargs(kw)
    model = dpro.model_constants("jp", n_segments=5, **mk)
    cols = dpro.apparatus_columns("jp", B, **ak)
    t2 = time.perf_counter()
    ev.simulate_host_cols(model, cols, B)
    t3 = time.perf_counter()
    T["total"].append(t1 - t0); T["columns+model"].append(t2 - t1); T["c_call"].append(t3 - t2)
    T["kernel_span"].append(ev.last_kernel_ms() * 1e-3)
for k, v in T.items():
    print(f"{k:16s} {np.median(v) * 1e3:8.3f} ms")
