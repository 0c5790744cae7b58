"""End-to-end rate of simulate_CZ_gate_batch (host arrays in, host arrays out) against the chunk size
of the pipelined C entry.  Usage: python tools/e2e_sweep.py [candidates]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from noisyquantumsimulator_b200 import rydberg_gates as rg  # noqa: E402
from noisyquantumsimulator_b200.engine import get_evaluator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
from noisyquantumsimulator_b200.rydberg_gates.simulation import METRIC_OUTPUTS  # noqa: E402
ev = get_evaluator(0)
# like bench.py's e2e leg: candidates in page-locked memory, the optimiser's metric columns back
cand = {k: ev.pinned_copy("c_" + k, v) for k, v in bench.make_candidates(B, 1234).items()}
for chunk in (65536, 32768, 24576, 20480, 16384, 12288):
    ev.set_chunk(chunk)
    for _ in range(2):
        rg.simulate_CZ_gate_batch("jp", B=B, prologue="device", copy=False, outputs=METRIC_OUTPUTS, **bench.MEDIUM, **cand)
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        r = rg.simulate_CZ_gate_batch("jp", B=B, prologue="device", copy=False, outputs=METRIC_OUTPUTS, **bench.MEDIUM, **cand)
    dt = (time.perf_counter() - t0) / n
    print(f"chunk {chunk:>10d}: {dt * 1e3:8.2f} ms/step  {B / dt / 1e6:7.2f} M evals/s  pipeline span {ev.last_kernel_ms():.2f} ms",
          flush=True)
