"""Batched CZ-gate evaluation on one B200 through the C ABI (include/rydcz.h).

``Evaluator`` is the thin host layer between the reference-facing Python API
(``rydberg_gates.simulation``) and the CUDA kernels.  torch is used for device memory,
streams and (in ``sharding``) torch.distributed; the arithmetic is all in csrc/.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib

OUT = dict(avg_fidelity=0, f00=1, f01=2, f10=3, f11=4, f11_pop=5, cz_phase_fidelity=6,
           controlled_phase=7, phase_error=8, phi01=9, phi10=10, phi11=11, loss=12, rydberg=13,
           steps=14, flops=15)


class RydczError(RuntimeError):
    pass


CUDA_STREAM_LEGACY = 1      # cudaStreamLegacy: the explicit handle of the default stream


def _torch_stream(device, stream=None) -> int:
    """Stream handle for the C ABI.  torch's default stream has the handle 0, which the C ABI reads as "the
    context's own (non-blocking) stream": pass cudaStreamLegacy instead, so that the kernels are ordered with the
    torch operations the caller enqueues on the default stream before and after the call."""
    import torch
    h = stream if stream is not None else torch.cuda.current_stream(device).cuda_stream
    return int(h) if h else CUDA_STREAM_LEGACY


def _host(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a, a.ctypes.data


class Evaluator:
    """One CUDA context-bound evaluator (one per process and GPU)."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        self.device = int(device)
        ctx = C.c_void_p()
        rc = self.lib.rydcz_create(self.device, C.byref(ctx))
        if rc != 0:
            msg = self.lib.rydcz_last_error(None)
            raise RydczError(f"rydcz_create(device={device}) failed: "
                             f"{msg.decode() if msg else rc} -- a CUDA GPU is required")
        self.ctx = ctx
        self._pinned = {}
        self._retired = []      # outgrown page-locked buffers: callers may still hold views of them

    def pinned(self, key: str, shape, dtype=np.float64) -> np.ndarray:
        """A page-locked NumPy array owned by this evaluator, reused across calls (grown on demand):
        staging for inputs and results of the *_host entry points."""
        dtype = np.dtype(dtype)
        need = int(np.prod(shape)) * dtype.itemsize
        buf = self._pinned.get(key)
        if buf is None or buf[1] < need:
            if buf is not None:
                # NumPy views of the old buffer (simulate_host*, copy=False results) may still be alive:
                # keep it until close() instead of freeing memory somebody can still read
                self._retired.append(buf[0])
            p = C.c_void_p()
            cap = max(need, 8)
            if self.lib.rydcz_pinned_alloc(cap, C.byref(p)) != 0:
                raise RydczError("cudaHostAlloc failed")
            buf = (p, cap)
            self._pinned[key] = buf
        raw = (C.c_char * buf[1]).from_address(buf[0].value)
        return np.frombuffer(raw, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def pinned_copy(self, key: str, array) -> np.ndarray:
        """A page-locked copy of `array` (same shape and dtype), owned by this evaluator.  Inputs that live in
        page-locked memory go to the device as they are: rydcz_simulate_host_cols / _select copy them with the H2D
        engine straight out of the caller's memory instead of staging them through the context's ring first."""
        a = np.ascontiguousarray(array)
        buf = self.pinned(key, a.shape, a.dtype)
        buf[...] = a
        return buf

    def close(self):
        for p, _ in getattr(self, "_pinned", {}).values():
            self.lib.rydcz_pinned_free(p)
        for p in getattr(self, "_retired", []):
            self.lib.rydcz_pinned_free(p)
        self._pinned, self._retired = {}, []
        if getattr(self, "ctx", None):
            self.lib.rydcz_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.rydcz_last_error(self.ctx)
            raise RydczError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def _check_batch(self, dim, points, segs, n_seg):
        """Shapes of one batch as the C ABI reads them (a wrong n_seg means out-of-bounds device reads)."""
        B = points.shape[0]
        assert points.ndim == 2 and points.shape[1] == self.lib.rydcz_point_stride_dim(dim), points.shape
        assert segs.ndim == 3 and segs.shape[0] == B and segs.shape[2] == _lib.SEG_STRIDE, (points.shape, segs.shape)
        if n_seg is not None:
            assert tuple(n_seg.shape) == (B,), n_seg.shape
            if B:
                lo, hi = int(n_seg.min()), int(n_seg.max())
                assert 0 <= lo and hi <= segs.shape[1], (lo, hi, segs.shape)
        return B

    # ------------------------------------------------------------------ host buffers
    def evaluate_host(self, dim: int, noisy: bool, points, segs, n_seg=None,
                      want_states: bool = False) -> Tuple[np.ndarray, Optional[np.ndarray]]:
        """points [B,12], segs [B,S,4], n_seg [B] int32 (None = all S).  Returns (out [B,16],
        states [B,stride] or None).  H2D, kernels and D2H all happen inside the C call."""
        points, p_ptr = _host(points, np.float64)
        segs, s_ptr = _host(segs, np.float64)
        n_ptr = None
        if n_seg is not None:
            n_seg, n_ptr = _host(n_seg, np.int32)
        B = self._check_batch(dim, points, segs, n_seg)
        out = np.empty((B, _lib.OUT_STRIDE))
        states = np.empty((B, self.lib.rydcz_state_stride(dim))) if want_states else None
        self._check(self.lib.rydcz_evaluate_host(
            self.ctx, dim, int(bool(noisy)), B, p_ptr, s_ptr, n_ptr, segs.shape[1],
            out.ctypes.data, states.ctypes.data if want_states else None), "rydcz_evaluate_host")
        return out, states

    def evaluate_shots_host(self, dim: int, noisy: bool, n_shots: int, seed: int, points, segs,
                            n_seg=None, want_shots: bool = False):
        points, p_ptr = _host(points, np.float64)
        segs, s_ptr = _host(segs, np.float64)
        n_ptr = None
        if n_seg is not None:
            n_seg, n_ptr = _host(n_seg, np.int32)
        B = self._check_batch(dim, points, segs, n_seg)
        assert n_shots >= 1
        mean = np.empty((B, _lib.OUT_STRIDE))
        shots = np.empty((B, n_shots, _lib.OUT_STRIDE)) if want_shots else None
        self._check(self.lib.rydcz_evaluate_shots_host(
            self.ctx, dim, int(bool(noisy)), B, int(n_shots), int(seed), p_ptr, s_ptr, n_ptr,
            segs.shape[1], mean.ctypes.data, shots.ctypes.data if want_shots else None),
            "rydcz_evaluate_shots_host")
        return mean, shots

    # ------------------------------------------------------------------ device tensors (torch)
    def evaluate_device(self, dim: int, noisy: bool, points, segs, n_seg=None, out=None,
                        states=None, want_states: bool = False, stream=None):
        """Same with CUDA torch tensors already in HBM; enqueues on the current torch stream and
        returns (out, states) tensors without synchronising."""
        import torch

        assert points.is_cuda and segs.is_cuda and points.dtype == torch.float64 and segs.dtype == torch.float64
        points, segs = points.contiguous(), segs.contiguous()
        if n_seg is not None:      # (values stay on the device: checking them would synchronise)
            assert n_seg.dtype == torch.int32 and n_seg.is_cuda and tuple(n_seg.shape) == (points.shape[0],)
        B = self._check_batch(dim, points, segs, None)
        if out is None:
            out = torch.empty((B, _lib.OUT_STRIDE), dtype=torch.float64, device=points.device)
        if want_states and states is None:
            states = torch.empty((B, self.lib.rydcz_state_stride(dim)), dtype=torch.float64,
                                 device=points.device)
        if n_seg is not None:
            n_seg = n_seg.contiguous()
            assert n_seg.dtype == torch.int32
        st = _torch_stream(points.device, stream)
        self._check(self.lib.rydcz_evaluate_device(
            self.ctx, dim, int(bool(noisy)), B, points.data_ptr(), segs.data_ptr(),
            n_seg.data_ptr() if n_seg is not None else None, segs.shape[1], out.data_ptr(),
            states.data_ptr() if states is not None else None, st), "rydcz_evaluate_device")
        return out, states

    def evaluate_shots_device(self, dim: int, noisy: bool, n_shots: int, seed: int, points, segs,
                              n_seg=None, mean=None, shots=None, want_shots=False, stream=None):
        import torch

        assert points.is_cuda and segs.is_cuda and points.dtype == torch.float64 and segs.dtype == torch.float64
        points, segs = points.contiguous(), segs.contiguous()
        if n_seg is not None:
            assert n_seg.dtype == torch.int32 and n_seg.is_cuda and tuple(n_seg.shape) == (points.shape[0],)
            n_seg = n_seg.contiguous()
        B = self._check_batch(dim, points, segs, None)
        assert n_shots >= 1
        if mean is None:
            mean = torch.empty((B, _lib.OUT_STRIDE), dtype=torch.float64, device=points.device)
        if want_shots and shots is None:
            shots = torch.empty((B, n_shots, _lib.OUT_STRIDE), dtype=torch.float64,
                                device=points.device)
        st = _torch_stream(points.device, stream)
        self._check(self.lib.rydcz_evaluate_shots_device(
            self.ctx, dim, int(bool(noisy)), B, int(n_shots), int(seed), points.data_ptr(),
            segs.data_ptr(), n_seg.data_ptr() if n_seg is not None else None, segs.shape[1],
            mean.data_ptr(), shots.data_ptr() if shots is not None else None, st),
            "rydcz_evaluate_shots_device")
        return mean, shots

    # ------------------------------------------------------------------ prologue on the device
    def simulate_host(self, model, app, n_shots: int = 0, seed: int = 1234, want_scalars: bool = True,
                      want_states: bool = False):
        """model [32], app [40,B] slot-major (rydberg_gates/device_prologue.py) -> (out [B,16], scalars [B,8] or
        None, states or None).  Prologue kernel, evaluation, copies: all inside the C call."""
        model, m_ptr = _host(model, np.float64)
        app, a_ptr = _host(app, np.float64)
        B = app.shape[1]
        assert model.shape == (_lib.MODEL_STRIDE,) and app.shape == (_lib.APP_STRIDE, B)
        # results land in page-locked buffers owned by the evaluator (valid until the next call)
        out = self.pinned("sim_out", (B, _lib.OUT_STRIDE))
        sc = self.pinned("sim_sc", (B, _lib.SCALAR_STRIDE)) if want_scalars else None
        st = self.pinned("sim_st", (B, self.lib.rydcz_state_stride(int(model[0])))) if want_states else None
        self._check(self.lib.rydcz_simulate_host(
            self.ctx, m_ptr, B, a_ptr, int(n_shots), int(seed), out.ctypes.data,
            sc.ctypes.data if want_scalars else None, st.ctypes.data if want_states else None),
            "rydcz_simulate_host")
        return out, sc, st

    def simulate_host_cols(self, model, cols, B: int, n_shots: int = 0, seed: int = 1234,
                           want_scalars: bool = True, want_states: bool = False):
        """Column form of ``simulate_host`` (rydcz_simulate_host_cols): ``cols`` is the list of 40
        apparatus slots from ``device_prologue.apparatus_columns`` -- floats for batch-wide constants,
        1-D float64 arrays (any stride) for slots that vary.  Only the varying slots are copied to
        the GPU, and the batch is pipelined in chunks over eight streams.  Results land in
        page-locked buffers owned by the evaluator (valid until the next call)."""
        model, m_ptr = _host(model, np.float64)
        assert model.shape == (_lib.MODEL_STRIDE,) and len(cols) == _lib.APP_STRIDE
        ptrs, strides, consts, keep = self._column_args(cols, B)
        out = self.pinned("sim_out", (B, _lib.OUT_STRIDE))
        sc = self.pinned("sim_sc", (B, _lib.SCALAR_STRIDE)) if want_scalars else None
        st = self.pinned("sim_st", (B, self.lib.rydcz_state_stride(int(model[0])))) if want_states else None
        self._check(self.lib.rydcz_simulate_host_cols(
            self.ctx, m_ptr, B, ptrs, strides, consts, int(n_shots), int(seed), out.ctypes.data,
            sc.ctypes.data if want_scalars else None, st.ctypes.data if want_states else None),
            "rydcz_simulate_host_cols")
        del keep
        return out, sc, st

    def _column_args(self, cols, B):
        """ctypes arrays (pointer, stride in doubles, batch-wide constant) for the 40 apparatus slots, plus the
        arrays that must stay alive during the call."""
        ptrs = (C.c_void_p * _lib.APP_STRIDE)()
        strides = (C.c_int64 * _lib.APP_STRIDE)()
        consts = (C.c_double * _lib.APP_STRIDE)()
        keep = []
        for k, v in enumerate(cols):
            if isinstance(v, np.ndarray):
                if v.dtype != np.float64 or v.ndim != 1 or v.strides[0] % 8 or v.strides[0] <= 0 \
                        or not v.flags.aligned:
                    v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1)
                assert v.shape == (B,)
                keep.append(v)
                ptrs[k], strides[k], consts[k] = v.ctypes.data, v.strides[0] // 8, np.nan
            else:
                ptrs[k], strides[k], consts[k] = None, 1, float(v)
        return ptrs, strides, consts, keep

    def simulate_host_select(self, model, cols, B: int, out_slots, scalar_slots=(), n_shots: int = 0,
                             seed: int = 1234):
        """``simulate_host_cols`` returning only the requested columns (rydcz_simulate_host_select):
        ``out_slots`` index the 16 result slots (``engine.OUT``), ``scalar_slots`` the 8 prologue scalars
        (``_lib.SCALARS``).  Returns two page-locked arrays, column-major: ``out[j]`` is the contiguous
        column of ``out_slots[j]`` over the B points (valid until the next call)."""
        model, m_ptr = _host(model, np.float64)
        assert model.shape == (_lib.MODEL_STRIDE,) and len(cols) == _lib.APP_STRIDE
        ptrs, strides, consts, keep = self._column_args(cols, B)
        o_idx = np.ascontiguousarray(out_slots, dtype=np.int32).reshape(-1)
        s_idx = np.ascontiguousarray(scalar_slots, dtype=np.int32).reshape(-1)
        out = self.pinned("sel_out", (max(len(o_idx), 1), B))
        sc = self.pinned("sel_sc", (max(len(s_idx), 1), B))
        self._check(self.lib.rydcz_simulate_host_select(
            self.ctx, m_ptr, B, ptrs, strides, consts, int(n_shots), int(seed),
            o_idx.ctypes.data if len(o_idx) else None, len(o_idx), out.ctypes.data,
            s_idx.ctypes.data if len(s_idx) else None, len(s_idx), sc.ctypes.data),
            "rydcz_simulate_host_select")
        del keep
        return out[:len(o_idx)], sc[:len(s_idx)]

    def point_status(self, n: int) -> np.ndarray:
        """enum rydcz_status of the first n items of the last evaluation call (int32 array)."""
        st = np.empty(int(n), dtype=np.int32)
        self._check(self.lib.rydcz_point_status(self.ctx, int(n), st.ctypes.data), "rydcz_point_status")
        return st

    def call_flags(self):
        """(flags, unsupported_points) of the last evaluation call: RYDCZ_FLAG_* bits (_lib.FLAG_*)."""
        f, u = C.c_int(), C.c_int64()
        self._check(self.lib.rydcz_call_flags(self.ctx, C.byref(f), C.byref(u)), "rydcz_call_flags")
        return int(f.value), int(u.value)

    def set_point_base(self, first_point: int):
        """Index of this evaluator's first point in the caller's whole batch (Monte-Carlo shots are keyed by the
        global point index, so a rank evaluating the slice [lo, hi) draws the shots of the unsharded call)."""
        self._check(self.lib.rydcz_set_point_base(self.ctx, int(first_point)), "rydcz_set_point_base")

    def set_pack_threads(self, threads: int):
        """Host threads that gather the varying columns into pinned staging (rydcz_set_pack_threads)."""
        self._check(self.lib.rydcz_set_pack_threads(self.ctx, int(threads)), "rydcz_set_pack_threads")

    def set_host_wait(self, sleep: bool):
        """True: the pipelined host entry points sleep while the GPU works instead of spinning on a core
        (rydcz_set_host_wait; for several ranks on one host)."""
        self._check(self.lib.rydcz_set_host_wait(self.ctx, 1 if sleep else 0), "rydcz_set_host_wait")

    def bind_host_to_gpu(self) -> int:
        """Bind this process's calling thread (and threads created later) to the CPUs next to this GPU; returns
        the number of CPUs in the new mask (0: topology unknown, nothing changed).  One call per rank."""
        n = C.c_int()
        self._check(self.lib.rydcz_bind_host_to_gpu(self.ctx, C.byref(n)), "rydcz_bind_host_to_gpu")
        return int(n.value)

    def set_chunk(self, points_per_chunk: int):
        self._check(self.lib.rydcz_set_chunk(self.ctx, int(points_per_chunk)), "rydcz_set_chunk")

    def simulate_device(self, model, app, n_shots: int = 0, seed: int = 1234, out=None, scalars=None,
                        stream=None):
        """Same with the apparatus rows already in HBM (CUDA torch tensor); asynchronous."""
        import torch

        model, m_ptr = _host(model, np.float64)
        app = app.contiguous()
        B = app.shape[1]
        assert app.shape[0] == _lib.APP_STRIDE
        if out is None:
            out = torch.empty((B, _lib.OUT_STRIDE), dtype=torch.float64, device=app.device)
        st = _torch_stream(app.device, stream)
        self._check(self.lib.rydcz_simulate_device(
            self.ctx, m_ptr, B, app.data_ptr(), int(n_shots), int(seed), out.data_ptr(),
            scalars.data_ptr() if scalars is not None else None, None, st), "rydcz_simulate_device")
        return out

    def prologue_device(self, model, app, points, segs, n_seg, scalars=None, stream=None):
        import torch

        model, m_ptr = _host(model, np.float64)
        st = _torch_stream(app.device, stream)
        self._check(self.lib.rydcz_prologue_device(
            self.ctx, m_ptr, app.shape[1], app.data_ptr(), points.data_ptr(), segs.data_ptr(), n_seg.data_ptr(),
            scalars.data_ptr() if scalars is not None else None, st), "rydcz_prologue_device")

    # ------------------------------------------------------------------ measurement
    def set_sort(self, enable: bool):
        self._check(self.lib.rydcz_set_sort(self.ctx, int(enable)), "rydcz_set_sort")

    def set_cluster(self, mode: int):
        """Clustered-spectrum propagation: 0 = never, 1 = where cheaper (default), 2 = wherever eligible."""
        self._check(self.lib.rydcz_set_cluster(self.ctx, int(mode)), "rydcz_set_cluster")

    def set_table(self, mode: int):
        """Shared coefficient table of the clustered path: 0 = off (records per slot), 1 = entries generated once
        per call (default), 2 = kept across calls."""
        self._check(self.lib.rydcz_set_table(self.ctx, int(mode)), "rydcz_set_table")

    def table_stats(self):
        """(table entries generated, slots of clustered points that read the table) since the last call."""
        a, b = C.c_int64(), C.c_int64()
        self._check(self.lib.rydcz_table_stats(self.ctx, C.byref(a), C.byref(b)), "rydcz_table_stats")
        return int(a.value), int(b.value)

    def cluster_stats(self, with_records: bool = False):
        """(points planned onto the clustered path, points handed back to the plain path[, coefficient
        records written]) since the last call."""
        a, b, r = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self.lib.rydcz_cluster_stats(self.ctx, C.byref(a), C.byref(b), C.byref(r)), "rydcz_cluster_stats")
        return (int(a.value), int(b.value), int(r.value)) if with_records else (int(a.value), int(b.value))

    def repeat_stats(self) -> int:
        """Points evaluated on the repeated-segment path (one propagator for all segments) since the last call."""
        a = C.c_int64()
        self._check(self.lib.rydcz_repeat_stats(self.ctx, C.byref(a)), "rydcz_repeat_stats")
        return int(a.value)

    def last_coef_ms(self) -> float:
        ms = C.c_float()
        self._check(self.lib.rydcz_last_coef_ms(self.ctx, C.byref(ms)), "rydcz_last_coef_ms")
        return float(ms.value)

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        self._check(self.lib.rydcz_last_kernel_ms(self.ctx, C.byref(ms)), "rydcz_last_kernel_ms")
        return float(ms.value)

    def launch_count(self) -> int:
        n = C.c_int64()
        self._check(self.lib.rydcz_launch_count(self.ctx, C.byref(n)), "rydcz_launch_count")
        return int(n.value)

    def fp64_peak(self) -> Tuple[float, float]:
        """(DFMA TFLOP/s, DMMA TFLOP/s) measured on this GPU by register-resident microbenchmarks."""
        a, b = C.c_double(), C.c_double()
        self._check(self.lib.rydcz_fp64_peak(self.ctx, C.byref(a), C.byref(b)), "rydcz_fp64_peak")
        return float(a.value), float(b.value)

    def fp64_matvec_peak(self) -> float:
        """TFLOP/s of the hot loop's own matvec (folded 3-level rho_11) run bare at the kernel's occupancy."""
        a = C.c_double()
        self._check(self.lib.rydcz_fp64_matvec_peak(self.ctx, C.byref(a)), "rydcz_fp64_matvec_peak")
        return float(a.value)


_default = {}


def get_evaluator(device: int = 0) -> Evaluator:
    if device not in _default:
        _default[device] = Evaluator(device)
    return _default[device]
