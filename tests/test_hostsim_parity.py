"""The kernel arithmetic (csrc/rydcz_core.cuh + generated systems), compiled for the host by
tests/hostsim, against the golden finals and the oracle.  CPU only; validates the algorithm
before any GPU time is spent.  The product never loads this harness."""
import numpy as np
import pytest

import _cases as C
from oracle import rydcz_oracle as orc


@pytest.fixture(scope="module")
def hostsim():
    return C.load_hostsim()


@pytest.mark.parametrize("case", C.load("reference_cases.json"), ids=lambda c: c["name"])
def test_states_and_fidelity_match_golden(hostsim, case):
    noisy = case["rates"] is not None
    pts, seg, ns = C.pack_batch([case])
    out, st = hostsim(case["dim"], noisy, pts, seg, ns)
    gold = C.golden_finals(case)
    ops = C.states_to_operators(case["dim"], st[0])
    for lab in ("01", "10", "11"):
        if noisy:
            assert np.abs(ops["rho" + lab] - gold[lab]).max() < 1e-11      # north_star asks 1e-8
        else:
            assert np.abs(ops["coh" + lab][:, 0] - gold[lab]).max() < 1e-11
    res = case["result_coherence"] if noisy else case["result"]
    o = out[0]
    assert abs(o[C.OUT["avg"]] - res["avg_fidelity"]) < 1e-10                # north_star asks 1e-6
    for lab, slot in (("00", "f00"), ("01", "f01"), ("10", "f10"), ("11", "f11")):
        assert abs(o[C.OUT[slot]] - res["fidelities"][lab]) < 1e-10
    assert abs(o[C.OUT["czf"]] - res["phase_info"]["cz_phase_fidelity"]) < 1e-10
    assert abs(o[C.OUT["ctrl"]] - res["phase_info"]["controlled_phase_rad"]) < 1e-9


@pytest.mark.parametrize("case", C.load("known_answers.json"), ids=lambda c: c["name"])
def test_known_answers(hostsim, case):
    pts, seg, ns = C.pack_batch([case])
    out, _ = hostsim(3, False, pts, seg, ns, want_states=False)
    exp, tol = case["expected_printed"], case["tolerance_printed"]
    assert abs(out[0, C.OUT["avg"]] - exp["avg_fidelity"]) <= tol["avg_fidelity"]
    if "f11" in exp:
        assert abs(out[0, C.OUT["f11"]] - exp["f11"]) <= tol["f11"]
    if "controlled_phase_deg" in exp:
        assert abs(np.degrees(out[0, C.OUT["ctrl"]]) - exp["controlled_phase_deg"]) <= 6e-3


def test_noise_to_zero_reduces_to_pure_branch(hostsim):
    """Gauge-free epilogue: with all rates zero the Lindblad path equals the Schroedinger path."""
    case = next(c for c in C.load("reference_cases.json") if c["name"] == "vb_bb5_default")
    pts, seg, ns = C.pack_batch([case])
    a, _ = hostsim(3, False, pts, seg, ns)
    b, _ = hostsim(3, True, pts, seg, ns)
    assert np.abs(a[0, :12] - b[0, :12]).max() < 1e-12


def test_trace_preserved_and_loss_accounting(hostsim):
    case = next(c for c in C.load("reference_cases.json") if c["name"] == "lp_noisy")
    pts, seg, ns = C.pack_batch([case])
    out, st = hostsim(3, True, pts, seg, ns)
    ops = C.states_to_operators(3, st[0])
    for lab in ("01", "10", "11"):
        assert abs(np.trace(ops["rho" + lab]).real - 1.0) < 1e-12
        assert np.abs(ops["rho" + lab] - ops["rho" + lab].conj().T).max() == 0.0
    rho = ops["rho11"]
    has0 = [i for i in range(9) if i // 3 == 0 or i % 3 == 0]
    assert abs(out[0, C.OUT["loss"]] - sum(rho[i, i].real for i in has0)) < 1e-14


def test_strong_damping_substeps(hostsim):
    """gamma*tau >> 1 forces the sub-stepping branch; still matches expm."""
    V, d1 = 2 * np.pi * 40e6, 1e5
    rates = orc.NoiseRates(gamma_r=4e6, gamma_phi_laser=3e7, gamma_scatter_intermediate=2e6,
                           gamma_loss_background=5e6)
    segs = [orc.Segment(0.5e-6, 2 * np.pi * 3e6 * np.exp(0.4j), 2 * np.pi * 1e6),
            orc.Segment(0.3e-6, 2 * np.pi * 3e6 * np.exp(-1.4j), 2 * np.pi * 1e6)]
    fin = orc.propagate(3, segs, V, d1, rates)
    pt = C.pack_point(V, d1, rates.collapsed())
    seg = np.array([[[s.dt, s.Omega.real, s.Omega.imag, s.Delta] for s in segs]])
    out, st = hostsim(3, True, pt[None], seg, np.array([2], dtype=np.int32))
    ops = C.states_to_operators(3, st[0])
    for lab in ("01", "10", "11"):
        assert np.abs(ops["rho" + lab] - fin[lab]).max() < 1e-11


def test_unequal_atoms_detuning_offsets(hostsim):
    """Per-atom detuning offsets (Monte-Carlo Doppler shots) break the atom symmetry."""
    V, d1 = 2 * np.pi * 300e6, 2e5
    rates = orc.NoiseRates(gamma_r=7e3, gamma_phi_laser=4e4, gamma_scatter_intermediate=3e3)
    segs = orc.segments_lp(2 * np.pi * 4e6, 0.377 * 2 * np.pi * 4e6, 4.29 / (2 * np.pi * 4e6))
    off = (2 * np.pi * 0.4e6, -2 * np.pi * 0.9e6)
    fin = orc.propagate(3, segs, V, d1, rates, detuning_offsets=off)
    pt = C.pack_point(V, d1, rates.collapsed(), off=off)
    seg = np.array([[[s.dt, s.Omega.real, s.Omega.imag, s.Delta] for s in segs]])
    out, st = hostsim(3, True, pt[None], seg, np.array([2], dtype=np.int32))
    ops = C.states_to_operators(3, st[0])
    for lab in ("01", "10", "11"):
        assert np.abs(ops["rho" + lab] - fin[lab]).max() < 1e-11
    assert abs(out[0, C.OUT["f01"]] - out[0, C.OUT["f10"]]) > 1e-6


@pytest.mark.parametrize("dim", [3, 4])
def test_exchange_folded_path_equals_general_path(hostsim, dim):
    """Identical atoms take the exchange-folded systems (15/6 variables instead of 25/8); the
    general systems must give the same rho, coherences and fidelities."""
    cases = [c for c in C.load("reference_cases.json") if c["dim"] == dim and len(c["segments"]["dt"]) <= 300]
    for noisy in (True, False):
        sel = [c for c in cases if (c["rates"] is not None) == noisy]
        pts, seg, ns = C.pack_batch(sel)
        a, sa = hostsim(dim, noisy, pts, seg, ns, allow_sym=True)
        b, sb = hostsim(dim, noisy, pts, seg, ns, allow_sym=False)
        assert np.abs(sa - sb).max() < 1e-11
        assert np.abs(a[:, :14] - b[:, :14]).max() < 1e-10
        assert np.all(a[:, 15] < 0.62 * b[:, 15])          # well under two thirds of the flops
