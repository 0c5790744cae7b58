"""Shared helpers for the parity tests: golden loading, packing into the C-ABI layout."""
from __future__ import annotations

import json
import os
from functools import lru_cache

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

PT_STRIDE, SG_STRIDE, OUT_STRIDE = 12, 4, 16
OUT = dict(avg=0, f00=1, f01=2, f10=3, f11=4, f11_pop=5, czf=6, ctrl=7, err=8, phi01=9, phi10=10,
           phi11=11, loss=12, ryd=13, steps=14, flops=15)
STATE_STRIDE = {3: 51, 4: 64, 5: 129}
STATE_OFFSETS = {3: dict(rho11=0, rho01=25, rho10=30, coh11=35, coh01=43, coh10=47),
                 4: dict(rho11=0, rho01=36, rho10=42, coh11=48, coh01=56, coh10=60),
                 5: dict(rho11=0, rho01=81, rho10=90, coh11=99, coh01=117, coh10=123)}


@lru_cache(None)
def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)["cases"]


def collapsed_rates(r):
    """(Γ_r→1, Γ_r→0, γ_φ, γ_sc, γ_mJ) from the ten reference rates (same-operator terms add)."""
    if r is None:
        return (0.0,) * 5
    b = r.get("branching_1", 0.5)
    return (r["gamma_r"] * b,
            r["gamma_r"] * (1 - b) + r["gamma_bbr"] + r["gamma_loss_antitrap"]
            + r["gamma_loss_background"] + r["gamma_leakage"],
            r["gamma_phi_laser"] + r["gamma_phi_thermal"] + r["gamma_phi_zeeman"],
            r["gamma_scatter_intermediate"], r["mJ_leakage_rate"])


def pack_point(V, delta1, rates5=(0, 0, 0, 0, 0), off=(0.0, 0.0), sig_det=0.0, sig_relr=0.0):
    row = np.zeros(PT_STRIDE)
    row[0], row[1], row[2], row[3] = V, delta1, off[0], off[1]
    row[4:9] = rates5
    row[9], row[10] = sig_det, sig_relr
    return row


def pack_case(case):
    s = case["segments"]
    seg = np.stack([s["dt"], s["re"], s["im"], s["delta"]], axis=1).astype(np.float64)
    pt = pack_point(case["scalars"]["V"], case["scalars"]["delta_1"], collapsed_rates(case["rates"]))
    return pt, seg


def pack_batch(cases):
    pts, segs = zip(*(pack_case(c) for c in cases))
    S = max(len(s) for s in segs)
    seg = np.zeros((len(cases), S, SG_STRIDE))
    for i, s in enumerate(segs):
        seg[i, :len(s)] = s
    return np.stack(pts), seg, np.array([len(s) for s in segs], dtype=np.int32)


def golden_finals(case):
    n = case["dim"] ** 2
    out = {}
    for k, v in case["finals"].items():
        a = np.array(v[0]) + 1j * np.array(v[1])
        out[k] = a.reshape(n, n) if a.size == n * n else a.reshape(n)
    return out


@lru_cache(None)
def systems(dim):
    from noisyquantumsimulator_b200.codegen import model
    return model.standard_systems(dim)


def states_to_operators(dim, row):
    """raw reduced state row -> {'rho11': (d²×d²) complex, ..., 'coh11': ...}"""
    S = systems(dim)
    out = {}
    for key, off in STATE_OFFSETS[dim].items():
        sysm = S[key]
        out[key] = sysm.to_operator(row[off:off + sysm.n])
    return out


def load_hostsim(variant: str = "", defines=()):
    """Build (if needed) and load the host harness of csrc/rydcz_core.cuh.  ``variant`` / ``defines``: an A/B build
    of the same sources under another name, e.g. ("chebrho", ["-DRYDCZ_CLOSED_RHO01=0"])."""
    import ctypes
    import subprocess

    d = os.path.join(HERE, "hostsim")
    so = os.path.join(d, f"libhostsim{'_' + variant if variant else ''}.so")
    srcs = [os.path.join(d, "hostsim.cpp")] + [
        os.path.join(HERE, "..", "noisyquantumsimulator_b200", "csrc", f)
        for f in ("rydcz_core.cuh", "rydcz_closed.cuh", "rydcz_cluster.cuh", "rydcz_repeat.cuh", "rydcz_prologue.cuh", "rydcz_systems_d3.inc", "rydcz_systems_d4.inc",
                  "rydcz_systems_d5.inc")]
    def stale():
        return not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)

    if stale():
        import fcntl
        with open(os.path.join(d, ".build.lock"), "w") as lock:      # xdist workers: one builds, the others wait
            fcntl.flock(lock, fcntl.LOCK_EX)
            if stale():
                tmp = so + f".tmp{os.getpid()}"
                subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", *defines, "-o", tmp, srcs[0]])
                os.replace(tmp, so)
    lib = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.hostsim_evaluate.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_longlong, dp, dp,
                                     ctypes.POINTER(ctypes.c_int), ctypes.c_int, dp, dp]
    lib.hostsim_evaluate.restype = ctypes.c_int
    lib.hostsim_allow_sym.argtypes = [ctypes.c_int]
    lib.hostsim_set_cluster.argtypes = [ctypes.c_int]

    lib.hostsim_set_table.argtypes = [ctypes.c_int]
    lib.hostsim_table_generated.restype = ctypes.c_longlong

    def evaluate(dim, noisy, points, segs, n_seg, want_states=True, allow_sym=True, cluster=0, table=1):
        """cluster: 0 = plain Chebyshev path, 1 = clustered-spectrum path where the planner prefers it,
        2 = wherever it is eligible (csrc/rydcz_cluster.cuh).  table: 0 = every slot generates coefficient records
        of its own, 1 = shared table, entries generated once per call, 2 = kept across calls."""
        lib.hostsim_allow_sym(int(allow_sym))
        lib.hostsim_set_cluster(int(cluster))
        lib.hostsim_set_table(int(table))
        import ctypes as C
        points = np.ascontiguousarray(points, dtype=np.float64)
        segs = np.ascontiguousarray(segs, dtype=np.float64)
        n_seg = np.ascontiguousarray(n_seg, dtype=np.int32)
        B = points.shape[0]
        out = np.zeros((B, OUT_STRIDE))
        st = np.zeros((B, STATE_STRIDE[dim])) if want_states else None
        rc = lib.hostsim_evaluate(dim, int(noisy), B, points.ctypes.data_as(dp),
                                  segs.ctypes.data_as(dp),
                                  n_seg.ctypes.data_as(C.POINTER(C.c_int)), segs.shape[1],
                                  out.ctypes.data_as(dp),
                                  st.ctypes.data_as(dp) if want_states else None)
        assert rc == 0
        lib.hostsim_set_cluster(0)
        return out, st

    def cluster_stats():
        a, b = ctypes.c_longlong(0), ctypes.c_longlong(0)
        lib.hostsim_cluster_stats(ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    lib.hostsim_repeat_used.restype = ctypes.c_longlong
    evaluate.repeat_used = lambda: int(lib.hostsim_repeat_used())
    evaluate.table_generated = lambda: int(lib.hostsim_table_generated())
    evaluate.cluster_stats = cluster_stats
    evaluate.lib = lib
    return evaluate


def plan_kwargs(case):
    """keyword arguments of rydberg_gates.prologue.build_plan for a golden case."""
    api = case["api"]
    e, nz = api["excitation"], api["noise"]
    kw = {k: (tuple(v) if k.startswith("qubit") else v) for k, v in api["kwargs"].items()}
    kw.update(api["params"])
    kw.update(laser_1_power=e["laser_1"]["power"], laser_1_waist=e["laser_1"]["waist"],
              laser_2_power=e["laser_2"]["power"], laser_2_waist=e["laser_2"]["waist"], Delta_e=e["Delta_e"],
              laser_1_linewidth_hz=e["laser_1"]["linewidth_hz"], laser_2_linewidth_hz=e["laser_2"]["linewidth_hz"],
              polarization_purity_1=e["laser_1"]["polarization_purity"],
              polarization_purity_2=e["laser_2"]["polarization_purity"],
              counter_propagating=e["counter_propagating"],
              include_motional_dephasing=nz["include_motional_dephasing"],
              include_doppler_dephasing=nz["include_doppler_dephasing"],
              include_intensity_noise=nz["include_intensity_noise"],
              intensity_noise_frac=nz["intensity_noise_frac"])
    return api["protocol"], kw


def api_inputs(case):
    """(simulation_inputs dataclass, simulate_CZ_gate kwargs) for a golden case."""
    from noisyquantumsimulator_b200 import rydberg_gates as rg
    api = case["api"]
    e = api["excitation"]
    exc = rg.TwoPhotonExcitationConfig(laser_1=rg.LaserParameters(**e["laser_1"]),
                                       laser_2=rg.LaserParameters(**e["laser_2"]), Delta_e=e["Delta_e"],
                                       counter_propagating=e["counter_propagating"])
    cls = {"lp": rg.LPSimulationInputs, "jp": rg.JPSimulationInputs, "sjp": rg.SmoothJPSimulationInputs}[api["protocol"]]
    inputs = cls(excitation=exc, noise=rg.NoiseSourceConfig(**api["noise"]), **api["params"])
    kw = {k: (tuple(v) if k.startswith("qubit") else v) for k, v in api["kwargs"].items()}
    return inputs, kw


class HostsimEvaluator:
    """Stands in for engine.Evaluator in CPU tests of the Python API layer: same methods, arithmetic
    from tests/hostsim (the g++ build of csrc/rydcz_core.cuh).  TEST ONLY."""

    def __init__(self):
        self._ev = load_hostsim()

    def evaluate_host(self, dim, noisy, points, segs, n_seg=None, want_states=False):
        if n_seg is None:
            n_seg = np.full(len(points), segs.shape[1], dtype=np.int32)
        out, st = self._ev(dim, noisy, points, segs, n_seg, want_states=True)
        return out, (st if want_states else None)


def hostsim_prologue(model, app):
    """Host build of the device prologue: (points, segs, n_seg, scalars)."""
    import ctypes as Ct
    load_hostsim()
    lib = Ct.CDLL(os.path.join(HERE, "hostsim", "libhostsim.so"))
    dp = Ct.POINTER(Ct.c_double)
    lib.hostsim_prologue.argtypes = [dp, Ct.c_longlong, dp, dp, dp, Ct.POINTER(Ct.c_int), dp]
    model = np.ascontiguousarray(model, dtype=np.float64)
    app = np.ascontiguousarray(app, dtype=np.float64)
    assert app.shape[0] == 40, "apparatus block is slot-major [40, B]"
    B, S = app.shape[1], int(model[2])
    pts, segs = np.zeros((B, 20 if int(model[0]) == 5 else PT_STRIDE)), np.zeros((B, S, SG_STRIDE))
    ns, sc = np.zeros(B, dtype=np.int32), np.zeros((B, 8))
    rc = lib.hostsim_prologue(model.ctypes.data_as(dp), B, app.ctypes.data_as(dp), pts.ctypes.data_as(dp),
                              segs.ctypes.data_as(dp), ns.ctypes.data_as(Ct.POINTER(Ct.c_int)),
                              sc.ctypes.data_as(dp))
    assert rc == 0
    return pts, segs, ns, sc


def _hostsim_simulate_host(self, model, app, n_shots=0, seed=1234, want_scalars=True, want_states=False):
    assert n_shots == 0, "the host harness has no shot sampler"
    pts, segs, ns, sc = hostsim_prologue(model, app)
    out, st = self._ev(int(model[0]), bool(model[3]), pts, segs, ns, want_states=True)
    return out, (sc if want_scalars else None), (st if want_states else None)


HostsimEvaluator.simulate_host = _hostsim_simulate_host


def _hostsim_simulate_host_cols(self, model, cols, B, n_shots=0, seed=1234, want_scalars=True, want_states=False):
    app = np.empty((40, B))
    for k, v in enumerate(cols):
        app[k] = v
    return _hostsim_simulate_host(self, model, app, n_shots, seed, want_scalars, want_states)


HostsimEvaluator.simulate_host_cols = _hostsim_simulate_host_cols


def _hostsim_simulate_host_select(self, model, cols, B, out_slots, scalar_slots=(), n_shots=0, seed=1234):
    """Column-major selection, as rydcz_simulate_host_select returns it."""
    out, sc, _ = _hostsim_simulate_host_cols(self, model, cols, B, n_shots, seed, True, False)
    return (np.ascontiguousarray(out[:, list(out_slots)].T).reshape(len(out_slots), B),
            np.ascontiguousarray(sc[:, list(scalar_slots)].T).reshape(len(scalar_slots), B))


HostsimEvaluator.simulate_host_select = _hostsim_simulate_host_select


# ---- 5-level model (hilbert_space_dim = 5) -----------------------------------------------------
PT5_STRIDE = 20
ACT2 = [5 * a + b for a in (1, 2, 3) for b in (1, 2, 3)]     # both atoms in {1, e, r}
ACT01, ACT10 = [1, 2, 3], [5, 10, 15]


def pack_batch5(cases):
    """five_level_cases.json entries -> (points [B,20], segs [B,S,4], n_seg [B]) (ragged batch)."""
    S = max(len(c["segs"]) for c in cases)
    pts = np.array([c["point"] for c in cases], dtype=np.float64)
    seg = np.zeros((len(cases), S, SG_STRIDE))
    for i, c in enumerate(cases):
        seg[i, :len(c["segs"])] = c["segs"]
    return pts, seg, np.array([len(c["segs"]) for c in cases], dtype=np.int32)


def _cplx(v):
    return np.array(v[0]) + 1j * np.array(v[1])


def check_five_level_case(case, out_row, st_row, tol_rho=1e-10, tol_f=1e-9):
    """One 5-level result row + raw state row against the oracle fixture."""
    exp = case["expected"]
    ops = states_to_operators(5, st_row)
    if case["noisy"]:
        for key, idx in (("rho11", ACT2), ("rho01", ACT01), ("rho10", ACT10)):
            got = ops[key][np.ix_(idx, idx)]
            assert np.abs(got - _cplx(exp[key + "_active"])).max() < tol_rho, (case["name"], key)
        assert abs(out_row[OUT["loss"]] - exp["loss"]) < tol_f
    else:
        assert np.abs(ops["coh11"][ACT2, 0] - _cplx(exp["psi11_active"])).max() < tol_rho
        assert np.abs(ops["coh01"][ACT01, 0] - _cplx(exp["psi01_active"])).max() < tol_rho
    assert abs(out_row[OUT["avg"]] - exp["avg_fidelity"]) < tol_f, case["name"]
    for lab, slot in (("00", "f00"), ("01", "f01"), ("10", "f10"), ("11", "f11")):
        assert abs(out_row[OUT[slot]] - exp["fidelities"][lab]) < tol_f
    assert abs(out_row[OUT["czf"]] - exp["cz_phase_fidelity"]) < tol_f
    d = out_row[OUT["ctrl"]] - exp["controlled_phase_rad"]
    assert abs((d + np.pi) % (2 * np.pi) - np.pi) < 1e-8
