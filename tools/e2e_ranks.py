"""End-to-end time of simulate_CZ_gate_batch with N ranks calling at the same time (torchrun): where the time of
the slowest rank goes, for a few host-side settings.  Diagnostic for the N = 8 end-to-end scaling."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from noisyquantumsimulator_b200 import rydberg_gates as rg  # noqa: E402
from noisyquantumsimulator_b200.engine import get_evaluator  # noqa: E402
from noisyquantumsimulator_b200.rydberg_gates import device_prologue as dpro  # noqa: E402
from noisyquantumsimulator_b200.rydberg_gates.simulation import METRIC_OUTPUTS  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = 262144
cand = {k: get_evaluator(local).pinned_copy("c_" + k, v) for k, v in bench.make_candidates(B, 1234 + rank).items()} if os.environ.get("E2E_PINNED", "1") == "1" else bench.make_candidates(B, 1234 + rank)
kw = dict(bench.MEDIUM, **cand)
ev = get_evaluator(local)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def gmax(x):
    t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def measure(label, reps=6):
    for _ in range(2):
        rg.simulate_CZ_gate_batch("jp", B=B, device=local, prologue="device", copy=False, outputs=METRIC_OUTPUTS, **kw)
    T = {"api": [], "py_cols": [], "c_call": [], "span": []}
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        rg.simulate_CZ_gate_batch("jp", B=B, device=local, prologue="device", copy=False, outputs=METRIC_OUTPUTS, **kw)
        t1 = time.perf_counter()
        barrier()
        mk, ak = dpro.split_kwargs(kw)
        model = dpro.model_constants("jp", n_segments=5, **mk)
        cols = dpro.apparatus_columns("jp", B, **ak)
        t2 = time.perf_counter()
        barrier()
        t3 = time.perf_counter()
        ev.simulate_host_select(model, cols, B, list(range(11)), [])
        t4 = time.perf_counter()
        T["api"].append(t1 - t0); T["py_cols"].append(t2 - t1); T["c_call"].append(t4 - t3)
        T["span"].append(ev.last_kernel_ms() * 1e-3)
    line = {k: round(gmax(float(np.median(v))) * 1e3, 3) for k, v in T.items()}
    if rank == 0:
        print(label, "max over ranks, ms:", line, "->", round(world * B / line["api"] / 1e3, 1), "M evals/s", flush=True)


if rank == 0:
    print("world", world, "cpus allowed", len(os.sched_getaffinity(0)), "of", os.cpu_count(), flush=True)
measure("default (pool)")
ev.set_host_wait(True)
measure("host wait = sleep")
ev.set_host_wait(False)
measure("host wait = spin")
if os.environ.get("E2E_QUICK") == "1":       # the N = 8 run: GPU-minutes are charged eight-fold
    ev.set_host_wait(True)
    n = ev.bind_host_to_gpu()
    measure(f"host wait = sleep, bound to {n} cpus")
    ev.set_host_wait(False)
    for chunk in (65536, 131072, 32768):
        ev.set_chunk(chunk)
        measure(f"spin, chunk={chunk}")
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0)
ev.set_pack_threads(1)
measure("pack_threads=1")
ev.set_pack_threads(2)
measure("pack_threads=2")
ev.set_pack_threads(1)
ev.set_chunk(32768)
measure("pack_threads=1 chunk=32768")
ev.set_chunk(131072)
measure("pack_threads=1 chunk=131072")
n = ev.bind_host_to_gpu()
if rank == 0:
    print("bound to", n, "cpus", flush=True)
ev.set_chunk(32768)
measure("bound, pack_threads=1")
if world > 1:
    dist.destroy_process_group()
