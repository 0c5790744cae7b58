#!/usr/bin/env python3
"""Do two front-end kernels run faster side by side?  The scalar prologue of one batch on one stream against the same
batch cut in 2 / 4 pieces on as many streams (two / four contexts).  If the pieces overlap, the front end of a
device-resident call (prologue + planner, 0.22 ms of a 1.77 ms step) could be issued in pieces."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from noisyquantumsimulator_b200.engine import Evaluator  # noqa: E402
from noisyquantumsimulator_b200.rydberg_gates import device_prologue as dpro  # noqa: E402

B = bench.B_PER_GPU
dev = torch.device("cuda", 0)
mk, ak = dpro.split_kwargs(dict(bench.MEDIUM, **bench.make_candidates(B, 1234)))
model = dpro.model_constants("jp", n_segments=bench.N_SEG, **mk)
app = torch.from_numpy(dpro.pack_apparatus("jp", B, **ak)).to(dev)
P = 4
evs = [Evaluator(0) for _ in range(P)]
streams = [torch.cuda.Stream() for _ in range(P)]


def run(pieces, iters=20):
    n = B // pieces
    apps = [app[:, k * n:(k + 1) * n].contiguous() for k in range(pieces)]
    pts = [torch.empty((n, 12), dtype=torch.float64, device=dev) for _ in range(pieces)]
    sgs = [torch.empty((n, bench.N_SEG, 4), dtype=torch.float64, device=dev) for _ in range(pieces)]
    nss = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(pieces)]

    def one():
        for k in range(pieces):
            with torch.cuda.stream(streams[k]):
                evs[k].prologue_device(model, apps[k], pts[k], sgs[k], nss[k])
    for _ in range(3):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.default_stream())
    for s in streams:
        s.wait_event(e0)
    for _ in range(iters):
        one()
    for s in streams:
        torch.cuda.default_stream().wait_stream(s)
    e1.record(torch.cuda.default_stream())
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for pieces in (1, 2, 4, 1, 2, 4):
    print(f"prologue of {B} points in {pieces} piece(s) on {pieces} stream(s): {run(pieces) * 1e3:.1f} us", flush=True)
