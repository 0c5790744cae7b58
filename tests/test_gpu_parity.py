"""Parity of the CUDA path (through the C ABI) with the oracle and the golden fixtures.
Everything here needs a B200; inputs are seeded, the oracle runs on the host beside it."""
import math

import numpy as np
import pytest

import _cases as C
from oracle import philox
from oracle import rydcz_oracle as orc

pytestmark = pytest.mark.gpu

REF = C.load("reference_cases.json")
KA = C.load("known_answers.json")


def _check_case(case, out_row, st_row):
    noisy = case["rates"] is not None
    gold = C.golden_finals(case)
    ops = C.states_to_operators(case["dim"], st_row)
    for lab in ("01", "10", "11"):
        if noisy:
            assert np.abs(ops["rho" + lab] - gold[lab]).max() < 1e-10      # north_star: 1e-8
        else:
            assert np.abs(ops["coh" + lab][:, 0] - gold[lab]).max() < 1e-10
    res = case["result_coherence"] if noisy else case["result"]
    assert abs(out_row[C.OUT["avg"]] - res["avg_fidelity"]) < 1e-9         # north_star: 1e-6
    for lab, slot in (("00", "f00"), ("01", "f01"), ("10", "f10"), ("11", "f11")):
        assert abs(out_row[C.OUT[slot]] - res["fidelities"][lab]) < 1e-9
    assert abs(out_row[C.OUT["czf"]] - res["phase_info"]["cz_phase_fidelity"]) < 1e-9


@pytest.mark.parametrize("dim,noisy", [(3, True), (3, False), (4, True), (4, False)])
def test_golden_cases_one_ragged_batch(evaluator, dim, noisy):
    """All fixtures of one (dim, branch) in ONE ragged batch: S from 2 to 998 side by side."""
    cases = [c for c in REF if c["dim"] == dim and (c["rates"] is not None) == noisy]
    pts, seg, ns = C.pack_batch(cases)
    out, st = evaluator.evaluate_host(dim, noisy, pts, seg, ns, want_states=True)
    for i, c in enumerate(cases):
        _check_case(c, out[i], st[i])


def test_known_answers_saved_qutip_outputs(evaluator):
    pts, seg, ns = C.pack_batch(KA)
    out, _ = evaluator.evaluate_host(3, False, pts, seg, ns)
    for i, case in enumerate(KA):
        exp, tol = case["expected_printed"], case["tolerance_printed"]
        assert abs(out[i, C.OUT["avg"]] - exp["avg_fidelity"]) <= tol["avg_fidelity"], case["name"]
        if "f11" in exp:
            assert abs(out[i, C.OUT["f11"]] - exp["f11"]) <= tol["f11"]
        if "cz_phase_fidelity" in exp:
            assert abs(out[i, C.OUT["czf"]] - exp["cz_phase_fidelity"]) <= tol["cz_phase_fidelity"]
        if "controlled_phase_deg" in exp:
            assert abs(math.degrees(out[i, C.OUT["ctrl"]]) - exp["controlled_phase_deg"]) <= 6e-3
        if "phase_error_deg" in exp:
            assert abs(math.degrees(out[i, C.OUT["err"]]) - exp["phase_error_deg"]) <= 6e-3


def _random_points(rng, n, protocol):
    """Seeded synthetic points in the ranges of SURVEY.md section 8d (configs 1, 2, 4)."""
    pts, segs = [], []
    for _ in range(n):
        Om = 2 * np.pi * rng.uniform(1.5e6, 9e6)
        V = Om * 10 ** rng.uniform(0.8, 2.4)
        d1 = 2 * np.pi * rng.uniform(0, 3e5)
        rates = orc.NoiseRates(gamma_r=rng.uniform(5e3, 9e3), gamma_phi_laser=rng.uniform(1e2, 5e4),
                               gamma_phi_thermal=rng.uniform(0, 5e4),
                               gamma_loss_antitrap=rng.uniform(0, 3e3), gamma_loss_background=1e3,
                               gamma_scatter_intermediate=rng.uniform(1e3, 2e4),
                               gamma_leakage=rng.uniform(0, 50))
        if protocol == "lp":
            sg = orc.segments_lp(Om, rng.uniform(0.2, 0.5) * Om, rng.uniform(3.5, 5.5) / Om)
        elif protocol == "bb":
            ot = rng.uniform(5, 40)
            sw = np.sort(rng.uniform(0.01, 0.99, 4)) * ot
            sg = orc.segments_bangbang(Om, ot, list(sw), list(rng.uniform(-np.pi, np.pi, 5)))
        else:
            sg = orc.segments_smooth_jp(Om, rng.uniform(5, 25) / Om, rng.uniform(0.05, 1) * np.pi,
                                        rng.uniform(0.5, 3) * Om, rng.uniform(0, 2 * np.pi),
                                        -rng.uniform(0.001, 0.1) * Om, 60)
        pts.append((V, d1, rates))
        segs.append(sg)
    return pts, segs


def _pack(pts, segs):
    S = max(len(s) for s in segs)
    P = np.stack([C.pack_point(V, d1, r.collapsed()) for V, d1, r in pts])
    G = np.zeros((len(pts), S, 4))
    for i, sg in enumerate(segs):
        G[i, :len(sg)] = [[s.dt, s.Omega.real, s.Omega.imag, s.Delta] for s in sg]
    return P, G, np.array([len(s) for s in segs], dtype=np.int32)


@pytest.mark.parametrize("protocol", ["lp", "bb", "sjp"])
@pytest.mark.parametrize("noisy", [True, False])
def test_random_batch_against_oracle(evaluator, protocol, noisy):
    rng = np.random.default_rng(1234)
    # (the dense oracle needs ~0.2 s per segment and point: 60-step smooth-JP points are kept few)
    pts, segs = _random_points(rng, 6 if protocol == "sjp" else 24, protocol)
    P, G, ns = _pack(pts, segs)
    out, st = evaluator.evaluate_host(3, noisy, P, G, ns, want_states=True)
    for i, ((V, d1, r), sg) in enumerate(zip(pts, segs)):
        fid, avg, info, fin = orc.evaluate(3, sg, V, d1, r if noisy else None, "exact", "coherence")
        ops = C.states_to_operators(3, st[i])
        for lab in ("01", "10", "11"):
            got = ops["rho" + lab] if noisy else ops["coh" + lab][:, 0]
            assert np.abs(got - fin[lab]).max() < 1e-10
        assert abs(out[i, C.OUT["avg"]] - avg) < 1e-9
        assert abs(out[i, C.OUT["f11"]] - fid["11"]) < 1e-9
        assert abs(out[i, C.OUT["czf"]] - info["cz_phase_fidelity"]) < 1e-9


def test_single_atom_rho_closed_form_and_its_fallback_in_one_warp(evaluator):
    """rho_01 / rho_10 come from the closed form (csrc/rydcz_closed.cuh) where it is accepted and from the Chebyshev
    loop where it declines (damping comparable to the drive; an idle segment on resonance).  Both kinds sit in the
    same warp here, identical and unequal atoms, and every state is compared with the expm oracle."""
    rng = np.random.default_rng(99)
    pts, segs = _random_points(rng, 16, "bb")
    for i, (V, d1, r) in enumerate(pts):
        if i % 2:                                    # dephasing of the order of the Rabi frequency: declined
            r.gamma_phi_laser = 0.4 * abs(segs[i][0].Omega)
    P, G, ns = _pack(pts, segs)
    P[8:, 2] = 2 * np.pi * rng.normal(0, 5e4, 8)      # unequal atoms: the general kernels, rho_01 != rho_10
    P[8:, 3] = 2 * np.pi * rng.normal(0, 5e4, 8)
    G[3, 1, 1:] = 0.0                                # an idle segment (Omega = 0) ...
    P[3, 1] = 0.0                                    # ... on resonance: declined as well
    out, st = evaluator.evaluate_host(3, True, P, G, ns, want_states=True)
    steps = out[:, C.OUT["steps"]]
    assert steps[1::2].min() > 1.5 * steps[0:8:2].max()      # the declined points ran the single-atom loop
    for i, ((V, d1, r), sg) in enumerate(zip(pts, segs)):
        sg = [orc.Segment(s.dt, complex(G[i, k, 1], G[i, k, 2]), s.Delta) for k, s in enumerate(sg)]
        fid, avg, info, fin = orc.evaluate(3, sg, V, P[i, 1], r, "exact", "coherence",
                                           detuning_offsets=(P[i, 2], P[i, 3]))
        ops = C.states_to_operators(3, st[i])
        for lab in ("01", "10", "11"):
            assert np.abs(ops["rho" + lab] - fin[lab]).max() < 1e-10, (i, lab)
        assert abs(out[i, C.OUT["avg"]] - avg) < 1e-9


def test_empty_and_single_point_batches(evaluator):
    out, st = evaluator.evaluate_host(3, True, np.zeros((0, 12)), np.zeros((0, 2, 4)),
                                      np.zeros(0, dtype=np.int32), want_states=True)
    assert out.shape == (0, 16) and st.shape == (0, 51)
    case = next(c for c in REF if c["name"] == "lp_noisy")
    pts, seg, ns = C.pack_batch([case])
    out, st = evaluator.evaluate_host(3, True, pts, seg, ns, want_states=True)
    _check_case(case, out[0], st[0])
    # a point with zero segments is the identity: F = 1, controlled phase 0 -> cz phase fidelity 0
    out, _ = evaluator.evaluate_host(3, True, pts, seg, np.array([0], dtype=np.int32))
    assert out[0, C.OUT["f01"]] == 1.0 and out[0, C.OUT["f11_pop"]] == 1.0
    assert abs(out[0, C.OUT["czf"]]) < 1e-30 and out[0, C.OUT["steps"]] == 0


def test_large_batch_properties_and_order_independence(evaluator):
    """65,536 points (config 1 shape): results do not depend on batch position, sorting on/off is
    bit-identical, rho stays a state (trace 1), replicated points agree bit for bit."""
    case = next(c for c in REF if c["name"] == "lp_noisy")
    pt, seg = C.pack_case(case)
    B = 65536
    rng = np.random.default_rng(7)
    P = np.repeat(pt[None], B, 0)
    G = np.repeat(seg[None], B, 0)
    scale = 10 ** rng.uniform(-0.5, 0.4, B)          # spread V (hence work) over an order of magnitude
    scale[:1024] = 1.0
    P[:, 0] *= scale
    out, st = evaluator.evaluate_host(3, True, P, G, None, want_states=True)
    assert np.all(out[:1024] == out[0])
    _check_case(case, out[0], st[0])
    tr = st[:, [0, 1, 4, 5, 8, 15, 20, 21, 24]].sum(1)   # diagonal variables of rho_11
    assert np.abs(tr - 1).max() < 1e-12
    assert np.all((out[:, 0] > 0.2) & (out[:, 0] <= 1.0))
    evaluator.set_sort(False)
    out2, _ = evaluator.evaluate_host(3, True, P, G, None)
    evaluator.set_sort(True)
    assert np.array_equal(out, out2)
    perm = rng.permutation(B)
    out3, _ = evaluator.evaluate_host(3, True, P[perm], G[perm], None)
    assert np.array_equal(out3, out[perm])


def test_device_entry_matches_host_entry(evaluator):
    import torch
    cases = [c for c in REF if c["dim"] == 3 and c["rates"] is not None][:12]
    pts, seg, ns = C.pack_batch(cases)
    out_h, st_h = evaluator.evaluate_host(3, True, pts, seg, ns, want_states=True)
    dev = torch.device("cuda:0")
    out_d, st_d = evaluator.evaluate_device(3, True, torch.from_numpy(pts).to(dev),
                                            torch.from_numpy(seg).to(dev),
                                            torch.from_numpy(ns).to(dev), want_states=True)
    torch.cuda.synchronize()
    assert np.array_equal(out_d.cpu().numpy(), out_h)
    assert np.array_equal(st_d.cpu().numpy(), st_h)
    assert evaluator.last_kernel_ms() > 0 and evaluator.launch_count() > 0


def test_shots_match_oracle_sampler_and_reduce(evaluator):
    """Monte-Carlo shots: every shot equals the oracle evaluated at the oracle-sampled parameters;
    the device mean equals the mean of the shots; sigma -> 0 reproduces the deterministic path."""
    case = next(c for c in REF if c["name"] == "lp_noisy_spacing5p4_weak")
    pt, seg = C.pack_case(case)
    P = np.repeat(pt[None], 3, 0)
    P[:, 9] = [0.0, 2 * np.pi * 40e3, 2 * np.pi * 150e3]    # sig_det
    P[:, 10] = [0.0, 0.004, 0.02]                          # sig_relr
    G = np.repeat(seg[None], 3, 0)
    n_shots, seed = 8, 1234
    mean, shots = evaluator.evaluate_shots_host(3, True, n_shots, seed, P, G, None, want_shots=True)
    det, _ = evaluator.evaluate_host(3, True, P[:1], G[:1], None)
    assert np.abs(shots[0, :, :14] - det[0, :14]).max() < 1e-13
    rates = orc.NoiseRates(**case["rates"])
    sg = [orc.Segment(*row) for row in zip(seg[:, 0], seg[:, 1] + 1j * seg[:, 2], seg[:, 3])]
    for i in (1, 2):
        for s in range(n_shots):
            oa, ob, Vs = philox.shot_parameters(seed, i, s, P[i, 0], P[i, 9], P[i, 10])
            fid, avg, info, _ = orc.evaluate(3, sg, Vs, P[i, 1], rates, "exact", "coherence", (oa, ob))
            assert abs(shots[i, s, C.OUT["avg"]] - avg) < 1e-9
            assert abs(shots[i, s, C.OUT["f11_pop"]] - info["F11_population"]) < 1e-9
    assert np.abs(mean[:, :14] - shots[:, :, :14].mean(1)).max() < 1e-14
    assert np.allclose(mean[:, 14], shots[:, :, 14].sum(1))


def test_fp64_peak_microbenchmark_is_sane(evaluator):
    dfma, dmma = evaluator.fp64_peak()
    assert 5.0 < dfma < 100.0 and dmma > 1.0
