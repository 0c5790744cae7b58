#!/usr/bin/env python3
"""Does planning of one batch overlap the evaluation of another?  Two contexts on two streams against one context
on one stream, same device-resident bang-bang batches (bench workload)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from noisyquantumsimulator_b200.engine import Evaluator  # noqa: E402
from noisyquantumsimulator_b200.rydberg_gates import device_prologue as dpro  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else bench.B_PER_GPU
dev = torch.device("cuda", 0)
apps = []
for j in range(4):
    mk, ak = dpro.split_kwargs(dict(bench.MEDIUM, **bench.make_candidates(B, 1234 + j)))
    apps.append(torch.from_numpy(dpro.pack_apparatus("jp", B, **ak)).to(dev))
model = dpro.model_constants("jp", n_segments=bench.N_SEG, **mk)
evs = [Evaluator(0), Evaluator(0)]
streams = [torch.cuda.Stream(), torch.cuda.Stream(priority=-1)]
outs = [torch.empty((B, 16), dtype=torch.float64, device=dev) for _ in range(2)]


def run(n_ctx, iters=12, split=1):
    def one(i):
        k = i % n_ctx
        with torch.cuda.stream(streams[k]):
            if split == 1:
                evs[k].simulate_device(model, apps[i % 4], out=outs[k])
            else:
                raise NotImplementedError
    for i in range(4):
        one(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.default_stream())
    for s in streams:
        s.wait_event(e0)
    for i in range(iters):
        one(i)
    for s in streams:
        torch.cuda.default_stream().wait_stream(s)
    e1.record(torch.cuda.default_stream())
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for n_ctx in (1, 2, 1, 2):
    ms = run(n_ctx)
    print(f"{n_ctx} context(s): {ms:.3f} ms per batch of {B}  ({B / ms / 1e3:.1f} M evals/s)", flush=True)
