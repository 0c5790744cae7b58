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


# ------------------------------------------------------------------------------------------------------------
# Sharded form of simulate_CZ_gate_batch: a DE generation (BASELINE.json configs[3]: 8192 candidates per
# iteration across 8 GPUs) or a sweep with Monte-Carlo shots (configs[2]) spread over the ranks of one box.
# ------------------------------------------------------------------------------------------------------------
def block_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice of rank `rank`: candidates of one generation (and shots of one sweep) cost about the
    same, so equal blocks balance; every rank derives the same partition without communication."""
    per = (B + world - 1) // world
    return min(B, rank * per), min(B, (rank + 1) * per)


def _slice_kwargs(kwargs: dict, B: int, lo: int, hi: int) -> dict:
    out = {}
    for k, v in kwargs.items():
        if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == B and B > 1:
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def simulate_sharded(protocol: str, B: int, *, rank: int = 0, world: int = 1, group=None, device: int = 0,
                     n_shots: int = 0, seed: int = 1234, outputs=None, evaluator=None, **kwargs):
    """``simulate_CZ_gate_batch(protocol, B=B, outputs=..., **kwargs)`` evaluated by ``world`` ranks.

    Every rank holds the (small) full input, evaluates the block ``block_range(B, rank, world)`` on its own GPU
    and contributes the requested result columns to ONE all-gather; every rank returns the full ``BatchResult``
    (``columns`` only).  With NCCL the block never leaves the device between the kernels and the collective:
    host columns -> H2D -> prologue + evaluate kernels -> column gather -> all_gather_into_tensor -> one D2H of
    the gathered columns.  Monte-Carlo shots are keyed by the global point index (``rydcz_set_point_base``), so
    the result equals the single-rank call bit for bit.  A rank whose evaluation raises contributes NaN rows and
    still joins the collective (the others must not hang); the exception is re-raised after it when
    ``world == 1``."""
    from .engine import get_evaluator
    from .rydberg_gates import simulation as sim

    outputs = tuple(outputs) if outputs is not None else sim.METRIC_OUTPUTS
    o_names = [o for o in outputs if o in sim.OUTPUT_SLOTS]
    s_names = [o for o in outputs if o in sim.SCALAR_SLOTS]
    if len(o_names) + len(s_names) != len(outputs):
        raise ValueError(f"unknown output column in {outputs}")
    names = o_names + s_names
    ev = evaluator if evaluator is not None else get_evaluator(device)
    kwargs = sim.shot_kwargs(kwargs, n_shots)
    if world == 1:
        return sim.simulate_CZ_gate_batch(protocol, B=B, n_shots=n_shots, seed=seed, device=device, outputs=outputs,
                                          **kwargs)
    import torch
    import torch.distributed as dist

    nccl = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", device) if nccl else torch.device("cpu")
    lo, hi = block_range(B, rank, world)
    per = (B + world - 1) // world
    mine = torch.full((len(names), per), float("nan"), dtype=torch.float64, device=dev)
    failure = None
    if hi > lo:
        try:
            kw = _slice_kwargs(kwargs, B, lo, hi)
            kw.pop("warn", None)
            model, cols = sim.device_prologue_inputs(protocol, hi - lo, kw)
            o_idx = [sim.OUTPUT_SLOTS[o] for o in o_names]
            s_idx = [sim.SCALAR_SLOTS[o] for o in s_names]
            if hasattr(ev, "set_point_base"):
                ev.set_point_base(lo)
            try:
                if nccl and hasattr(ev, "simulate_device"):
                    app = np.empty((len(cols), hi - lo))
                    for k, v in enumerate(cols):
                        app[k] = v
                    d_app = torch.from_numpy(app).to(dev, non_blocking=True)
                    sc = torch.empty((hi - lo, 8), dtype=torch.float64, device=dev)
                    out = ev.simulate_device(model, d_app, n_shots, seed, scalars=sc)
                    if o_idx:
                        mine[:len(o_idx), :hi - lo] = out[:, o_idx].T
                    if s_idx:
                        mine[len(o_idx):, :hi - lo] = sc[:, s_idx].T
                else:
                    o_cm, s_cm = ev.simulate_host_select(model, cols, hi - lo, o_idx, s_idx, n_shots, seed)
                    if o_idx:
                        mine[:len(o_idx), :hi - lo] = torch.from_numpy(np.array(o_cm)).to(dev)
                    if s_idx:
                        mine[len(o_idx):, :hi - lo] = torch.from_numpy(np.array(s_cm)).to(dev)
            finally:
                if hasattr(ev, "set_point_base"):
                    ev.set_point_base(0)
        except Exception as e:          # contribute NaN rows, stay in the collective
            failure = e
            mine.fill_(float("nan"))
    full = torch.empty((world * len(names), per), dtype=torch.float64, device=dev)     # rank-major concatenation
    dist.all_gather_into_tensor(full, mine, group=group)
    host = full.cpu().numpy().reshape(world, len(names), per)
    columns = {}
    for j, name in enumerate(names):
        columns[name] = np.concatenate([host[r, j, :max(0, min(B, (r + 1) * per) - r * per)] for r in range(world)])
    g = columns.get
    res = sim.BatchResult(avg_fidelity=g("avg_fidelity"), f11_population=g("f11_population"),
                          cz_phase_fidelity=g("cz_phase_fidelity"), controlled_phase_rad=g("controlled_phase_rad"),
                          phase_error_from_pi_rad=g("phase_error_from_pi_rad"), loss_probability=g("loss_probability"),
                          rydberg_residual=g("rydberg_residual"), tau_total=g("tau_total"), Omega=g("Omega"), V=g("V"),
                          columns=columns)
    res.local_failure = failure
    return res
