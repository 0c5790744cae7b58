"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/rydcz.h
declares.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from noisyquantumsimulator_b200 import build, _lib
    build.build(verbose=False)
    return _lib.load()


def test_header_and_binding_agree(lib):
    from noisyquantumsimulator_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "rydcz.h")).read()
    declared = set(re.findall(r"\b(rydcz_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.exported_symbols())
    for name in declared:
        assert hasattr(lib, name), name


def test_layout_queries(lib):
    assert lib.rydcz_version() >= 100
    assert (lib.rydcz_point_stride(), lib.rydcz_seg_stride(), lib.rydcz_out_stride()) == (12, 4, 16)
    assert lib.rydcz_state_stride(3) == 51 and lib.rydcz_state_stride(4) == 64
    assert lib.rydcz_state_stride(5) == 129 and lib.rydcz_state_stride(6) == -1
    assert [lib.rydcz_point_stride_dim(d) for d in (3, 4, 5, 6)] == [12, 12, 20, -1]
    import _cases as C
    for dim in (3, 4, 5):
        for name, off in C.STATE_OFFSETS[dim].items():
            assert lib.rydcz_state_offset(dim, name.encode()) == off
            assert lib.rydcz_system_size(dim, name.encode()) == C.systems(dim)[name].n
    assert lib.rydcz_system_nnz(3, b"rho11") == 169
    assert lib.rydcz_state_offset(3, b"nope") == -1


def test_generated_sources_are_current():
    """csrc/*.inc is what codegen/emit.py prints today."""
    from noisyquantumsimulator_b200.codegen import emit
    for d in (3, 4, 5):
        path = os.path.join(ROOT, "noisyquantumsimulator_b200", "csrc", f"rydcz_systems_d{d}.inc")
        assert open(path).read() == emit.emit_dim(d)


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from noisyquantumsimulator_b200.engine import Evaluator, RydczError
    with pytest.raises(RydczError):
        Evaluator(0)


def test_torch_ops_are_registered_and_have_no_cpu_kernel(lib):
    """torch.ops.rydcz.* (csrc/rydcz_torch.cpp): registered in the dispatcher, shapes propagate through the Meta
    kernels, and CPU tensors are refused -- there is no CPU path behind the operators."""
    import torch
    from noisyquantumsimulator_b200 import build
    build.build_torch_extension(verbose=False)
    from noisyquantumsimulator_b200 import torch_ops  # noqa: F401
    for name in ("evaluate", "evaluate_shots", "simulate", "point_status"):
        assert hasattr(torch.ops.rydcz, name)
    p = torch.empty(7, 12, dtype=torch.float64, device="meta")
    s = torch.empty(7, 5, 4, dtype=torch.float64, device="meta")
    out, st = torch.ops.rydcz.evaluate(p, s, None, 3, True, True)
    assert out.shape == (7, 16) and st.shape == (7, 51)
    out, st = torch.ops.rydcz.evaluate(torch.empty(7, 20, dtype=torch.float64, device="meta"), s, None, 5, True, False)
    assert out.shape == (7, 16) and st.shape == (0, 129)
    mean, shots = torch.ops.rydcz.evaluate_shots(p, s, None, 3, True, 8, 1234, True)
    assert mean.shape == (7, 16) and shots.shape == (7, 8, 16)
    out, sc = torch.ops.rydcz.simulate(torch.empty(32, dtype=torch.float64, device="meta"),
                                       torch.empty(40, 9, dtype=torch.float64, device="meta"), 0, 1234)
    assert out.shape == (9, 16) and sc.shape == (9, 8)
    with pytest.raises(NotImplementedError):
        torch.ops.rydcz.evaluate(torch.zeros(2, 12, dtype=torch.float64), torch.zeros(2, 1, 4, dtype=torch.float64),
                                 None, 3, True, False)
