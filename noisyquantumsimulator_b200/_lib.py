"""ctypes binding of include/rydcz.h.  There is no CPU fallback: if the CUDA library is missing
or no GPU is visible, the product path raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# RYDCZ_LIB_PATH: load another build of the same library (A/B timing of kernel variants)
LIB_PATH = os.environ.get("RYDCZ_LIB_PATH") or os.path.join(HERE, "csrc", "librydcz.so")

POINT_STRIDE, SEG_STRIDE, OUT_STRIDE = 12, 4, 16
APP_STRIDE, MODEL_STRIDE, SCALAR_STRIDE = 40, 32, 8
SYSTEMS = ("rho11", "rho01", "rho10", "coh11", "coh01", "coh10")
# enum rydcz_status / RYDCZ_FLAG_* (include/rydcz.h)
STATUS = {0: "plain", 1: "plain_folded", 2: "cluster_general", 3: "cluster", 5: "repeat", 6: "unsupported"}
STATUS_HANDED_BACK = 0x10
FLAG_NO_CLUSTER_SCRATCH, FLAG_UNSUPPORTED_POINTS = 1, 2
# scalar columns the prologue kernel reports (csrc/rydcz_prologue.cuh: enum ScalarSlot)
SCALARS = dict(Omega=0, V=1, tau_total=2, tau_single=3, Delta=4, xi_re=5, xi_im=6, omega_tau=7)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p

_SIGNATURES = {
    "rydcz_version": (C.c_int, []),
    "rydcz_point_stride": (C.c_int, []),
    "rydcz_seg_stride": (C.c_int, []),
    "rydcz_out_stride": (C.c_int, []),
    "rydcz_point_stride_dim": (C.c_int, [C.c_int]),
    "rydcz_state_stride": (C.c_int, [C.c_int]),
    "rydcz_state_offset": (C.c_int, [C.c_int, C.c_char_p]),
    "rydcz_system_size": (C.c_int, [C.c_int, C.c_char_p]),
    "rydcz_system_nnz": (C.c_int, [C.c_int, C.c_char_p]),
    "rydcz_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "rydcz_destroy": (None, [_vp]),
    "rydcz_last_error": (C.c_char_p, [_vp]),
    "rydcz_set_sort": (C.c_int, [_vp, C.c_int]),
    "rydcz_set_cluster": (C.c_int, [_vp, C.c_int]),
    "rydcz_set_table": (C.c_int, [_vp, C.c_int]),
    "rydcz_table_stats": (C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "rydcz_set_point_base": (C.c_int, [_vp, C.c_int64]),
    "rydcz_cluster_stats": (C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "rydcz_evaluate_device": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp, C.c_int,
                                        _vp, _vp, _vp]),
    "rydcz_evaluate_host": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp, C.c_int,
                                      _vp, _vp]),
    "rydcz_evaluate_shots_device": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_uint64,
                                              _vp, _vp, _vp, C.c_int, _vp, _vp, _vp]),
    "rydcz_evaluate_shots_host": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_uint64,
                                            _vp, _vp, _vp, C.c_int, _vp, _vp]),
    "rydcz_app_stride": (C.c_int, []),
    "rydcz_model_stride": (C.c_int, []),
    "rydcz_scalar_stride": (C.c_int, []),
    "rydcz_prologue_device": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rydcz_simulate_device": (C.c_int, [_vp, _vp, C.c_int64, _vp, C.c_int, C.c_uint64, _vp, _vp, _vp, _vp]),
    "rydcz_simulate_host": (C.c_int, [_vp, _vp, C.c_int64, _vp, C.c_int, C.c_uint64, _vp, _vp, _vp]),
    "rydcz_simulate_host_cols": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, _vp, C.c_int, C.c_uint64, _vp, _vp, _vp]),
    "rydcz_simulate_host_select": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, _vp, C.c_int, C.c_uint64, _vp, C.c_int,
                                             _vp, _vp, C.c_int, _vp]),
    "rydcz_set_chunk": (C.c_int, [_vp, C.c_int64]),
    "rydcz_set_pack_threads": (C.c_int, [_vp, C.c_int]),
    "rydcz_set_host_wait": (C.c_int, [_vp, C.c_int]),
    "rydcz_bind_host_to_gpu": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "rydcz_point_status": (C.c_int, [_vp, C.c_int64, _vp]),
    "rydcz_call_flags": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "rydcz_pinned_alloc": (C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    "rydcz_pinned_free": (C.c_int, [_vp]),
    "rydcz_last_kernel_ms": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "rydcz_last_coef_ms": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "rydcz_repeat_stats": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "rydcz_launch_count": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "rydcz_fp64_peak": (C.c_int, [_vp, _dp, _dp]),
    "rydcz_fp64_matvec_peak": (C.c_int, [_vp, _dp]),
}

_lib = None


def exported_symbols():
    """Every entry point include/rydcz.h declares."""
    return tuple(_SIGNATURES)


def load():
    """dlopen csrc/librydcz.so and attach signatures.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m noisyquantumsimulator_b200.build` "
            "(or __graft_entry__.build()).  There is no CPU fallback for the CZ-gate path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    assert lib.rydcz_point_stride() == POINT_STRIDE and lib.rydcz_out_stride() == OUT_STRIDE
    _lib = lib
    return lib
