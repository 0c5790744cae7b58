"""Sharding of a parameter grid / optimiser population over the GPUs of one box.

Every parameter point is independent, so the flat point index is cut into ``world`` contiguous
blocks after a work-balancing interleave (points sorted by estimated work are dealt round-robin,
so each rank gets the same mix of cheap and expensive points).  Each rank evaluates its block on
its own GPU; the only exchange is ONE all-gather of the per-point output rows (16 doubles each:
8192 points -> 1 MiB), over NCCL/NVLink on GPUs and gloo in the CPU tests.  There is no
data-path collective inside the evolution.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

OUT_STRIDE = 16


def estimated_work(points: np.ndarray, segs: np.ndarray, n_seg: Optional[np.ndarray] = None) -> np.ndarray:
    """Same estimate the device pre-pass uses (csrc/rydcz_core.cuh: point_work)."""
    V = np.abs(points[:, 0])[:, None]
    damp = points[:, 4:8].sum(1)[:, None]
    R = (V + 2 * (np.abs(segs[:, :, 1]) + np.abs(segs[:, :, 2])) + 2 * np.abs(segs[:, :, 3])
         + np.abs(points[:, 2:3]) + np.abs(points[:, 3:4]) + 2 * np.abs(points[:, 1:2]) + 2 * damp)
    z = R * np.abs(segs[:, :, 0])
    w = np.where(z > 0, z + 10.3 * np.cbrt(z) + 5.0, 0.0)
    if n_seg is not None:
        w = np.where(np.arange(segs.shape[1])[None, :] < n_seg[:, None], w, 0.0)
    return w.sum(1)


def shard_assignment(work: np.ndarray, world: int) -> np.ndarray:
    """rank of every point: sort by work (descending, stable) and deal round-robin."""
    order = np.argsort(-work, kind="stable")
    owner = np.empty(len(work), dtype=np.int64)
    owner[order] = np.arange(len(work)) % world
    return owner


def local_indices(owner: np.ndarray, rank: int) -> np.ndarray:
    return np.nonzero(owner == rank)[0]


def evaluate_sharded(evaluator, dim: int, noisy: bool, points: np.ndarray, segs: np.ndarray,
                     n_seg: Optional[np.ndarray] = None, *, rank: int = 0, world: int = 1, group=None,
                     device=None) -> np.ndarray:
    """Evaluate the whole batch across ``world`` ranks; every rank returns the full [B,16] result.

    All ranks hold the full input (it is small: a population of 8192 candidates is ~1 MB), so only
    results travel.  ``evaluator`` is an engine.Evaluator bound to this rank's GPU.
    """
    B = points.shape[0]
    if world == 1:
        return evaluator.evaluate_host(dim, noisy, points, segs, n_seg)[0]
    import torch
    import torch.distributed as dist

    owner = shard_assignment(estimated_work(points, segs, n_seg), world)
    mine = local_indices(owner, rank)
    out_local, _ = evaluator.evaluate_host(dim, noisy, points[mine], segs[mine],
                                           None if n_seg is None else n_seg[mine])
    cap = (B + world - 1) // world              # round-robin deal => every rank has <= cap points
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                             if dist.get_backend(group) == "nccl" else torch.device("cpu"))
    send = torch.zeros((cap, OUT_STRIDE), dtype=torch.float64, device=dev)
    send[:len(mine)] = torch.from_numpy(out_local).to(dev)
    recv = torch.empty((world * cap, OUT_STRIDE), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().numpy().reshape(world, cap, OUT_STRIDE)
    out = np.empty((B, OUT_STRIDE))
    for r in range(world):
        idx = local_indices(owner, r)
        out[idx] = recv[r, :len(idx)]
    return out
