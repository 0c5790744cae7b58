"""``torch.ops.rydcz.*``: the evaluator as registered PyTorch operators (csrc/rydcz_torch.cpp).

    import noisyquantumsimulator_b200.torch_ops      # loads csrc/librydcz_torch.so, registers the ops
    out, states = torch.ops.rydcz.evaluate(points, segs, n_seg, 3, True, False)     # CUDA float64 tensors
    out, scalars = torch.ops.rydcz.simulate(model, app, 0, 1234)                    # device prologue + evaluate
    mean, shots = torch.ops.rydcz.evaluate_shots(points, segs, None, 3, True, 64, 1234, False)
    status = torch.ops.rydcz.point_status(B)                                         # enum rydcz_status per point

Work is enqueued on the current torch stream of the tensors' device and nothing synchronises.  There is no CPU
kernel behind these ops: CPU tensors raise ``NotImplementedError`` from the dispatcher.  The Meta kernels only
propagate shapes (tracing).  Output layout: ``engine.OUT`` / include/rydcz.h.
"""
from __future__ import annotations

import os

import torch

from . import build as _build

_LOADED = False


def load() -> None:
    """Load the extension once (idempotent).  Raises if it has not been built: run
    ``python -m noisyquantumsimulator_b200.build`` (or ``__graft_entry__.build()``)."""
    global _LOADED
    if _LOADED:
        return
    if not os.path.exists(_build.TORCH_LIB):
        raise RuntimeError(f"{_build.TORCH_LIB} is missing: build it with `python -m noisyquantumsimulator_b200.build`")
    torch.ops.load_library(_build.TORCH_LIB)
    _LOADED = True


load()

evaluate = torch.ops.rydcz.evaluate
evaluate_shots = torch.ops.rydcz.evaluate_shots
simulate = torch.ops.rydcz.simulate
point_status = torch.ops.rydcz.point_status

__all__ = ["load", "evaluate", "evaluate_shots", "simulate", "point_status"]
