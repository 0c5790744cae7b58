"""5-level model (25x25 rho, BASELINE.json configs[4]) on the CUDA path, through the C ABI."""
import numpy as np
import pytest

import _cases as C
from oracle import rydcz_oracle as orc

pytestmark = pytest.mark.gpu
CASES = C.load("five_level_cases.json")


@pytest.mark.parametrize("noisy", [True, False])
def test_oracle_fixtures_one_ragged_batch(evaluator, noisy):
    cases = [c for c in CASES if c["noisy"] == noisy]
    pts, seg, ns = C.pack_batch5(cases)
    out, st = evaluator.evaluate_host(5, noisy, pts, seg, ns, want_states=True)
    for i, c in enumerate(cases):
        C.check_five_level_case(c, out[i], st[i])


def test_random_batch_against_dense_oracle(evaluator):
    """Seeded Rb/Cs-like parameter points, LP and bang-bang, against the dense 625x625 expm."""
    rng = np.random.default_rng(55)
    pts, segs, rates = [], [], []
    for i in range(10):
        Om = 2 * np.pi * rng.uniform(1.5e6, 6e6)
        V = Om * 10 ** rng.uniform(0.8, 2.2)
        De = 2 * np.pi * rng.uniform(0.8e9, 6e9)
        om1 = np.sqrt(2 * De * Om) * rng.uniform(0.5, 2.0)
        r = orc.NoiseRates5(g_e1=rng.uniform(1e7, 2e7), g_e0=rng.uniform(1e7, 2e7), g_r1=rng.uniform(2e3, 5e3),
                            g_r0=rng.uniform(2e3, 5e3), g_rl=rng.uniform(0, 3e3), g_phi=rng.uniform(1e3, 1e5))
        if i % 2:
            sg = orc.segments_lp(Om, rng.uniform(0.2, 0.5) * Om, rng.uniform(3.5, 5.5) / Om)
        else:
            ot = rng.uniform(5, 20)
            sw = np.sort(rng.uniform(0.01, 0.99, 4)) * ot
            sg = orc.segments_bangbang(Om, ot, list(sw), list(rng.uniform(-np.pi, np.pi, 5)))
        row = np.zeros(C.PT5_STRIDE)
        row[[0, 1, 4, 5, 6, 12, 13, 14, 15, 16, 17]] = [V, 2 * np.pi * rng.uniform(0, 3e5), r.g_r1, r.g_r0, r.g_phi,
                                                        om1, De, r.g_e1, r.g_e0, r.g_rl, float(i % 3 != 0)]
        pts.append(row); segs.append(sg); rates.append(r)
    S = max(len(s) for s in segs)
    G = np.zeros((len(pts), S, 4))
    for i, sg in enumerate(segs):
        G[i, :len(sg)] = [[s.dt, s.Omega.real, s.Omega.imag, s.Delta] for s in sg]
    ns = np.array([len(s) for s in segs], dtype=np.int32)
    P = np.stack(pts)
    out, st = evaluator.evaluate_host(5, True, P, G, ns, want_states=True)
    for i, (row, sg, r) in enumerate(zip(pts, segs, rates)):
        fid, avg, info, fin = orc.evaluate5(sg, row[0], row[1], row[12], row[13], r, ls_comp=row[17])
        ops = C.states_to_operators(5, st[i])
        assert np.abs(ops["rho11"][np.ix_(C.ACT2, C.ACT2)] - fin["11"][np.ix_(C.ACT2, C.ACT2)]).max() < 1e-10
        assert np.abs(ops["rho01"][np.ix_(C.ACT01, C.ACT01)] - fin["01"][np.ix_(C.ACT01, C.ACT01)]).max() < 1e-10
        assert abs(out[i, C.OUT["avg"]] - avg) < 1e-9 and abs(out[i, C.OUT["f11"]] - fid["11"]) < 1e-9


def test_large_batch_properties(evaluator):
    """16,384 points: position independence, sorting on/off bit-identical, far-detuned points approach
    the 3-level kernel's answer."""
    case = next(c for c in CASES if c["name"] == "lp5_rb_noisy")
    B = 16384
    rng = np.random.default_rng(9)
    P = np.repeat(np.array(case["point"])[None], B, 0)
    G = np.repeat(np.array(case["segs"])[None], B, 0)
    scale = 10 ** rng.uniform(-0.4, 0.3, B)
    scale[:512] = 1.0
    P[:, 0] *= scale
    out, st = evaluator.evaluate_host(5, True, P, G, None, want_states=True)
    assert np.all(out[:512] == out[0])
    C.check_five_level_case(case, out[0], st[0])
    diag = [0, 17, 32, 45, 56, 65, 72, 77, 80]               # populations of the 9 active two-atom states
    assert np.abs(st[:, diag].sum(1) + out[:, C.OUT["loss"]] - 1).max() < 1e-12
    evaluator.set_sort(False)
    out2, _ = evaluator.evaluate_host(5, True, P, G, None)
    evaluator.set_sort(True)
    assert np.array_equal(out, out2)
    perm = rng.permutation(B)
    out3, _ = evaluator.evaluate_host(5, True, P[perm], G[perm], None)
    assert np.array_equal(out3, out[perm])


def _oracle5(case, pt, V=None, offs=None):
    segs = [orc.Segment(s[0], complex(s[1], s[2]), s[3]) for s in case["segs"]]
    rates = orc.NoiseRates5(g_e1=pt[14], g_e0=pt[15], g_r1=pt[4], g_r0=pt[5], g_rl=pt[16], g_phi=pt[6]) \
        if case["noisy"] else None
    offs = (pt[2], pt[3]) if offs is None else offs
    return orc.evaluate5(segs, pt[0] if V is None else V, pt[1], pt[12], pt[13], rates, ls_comp=pt[17],
                         detuning_offsets=offs)


@pytest.mark.parametrize("name", ["lp5_rb_noisy", "lp5_rb_noiseless", "lp5_cs_noisy"])
def test_unequal_atoms_general_systems_against_dense_oracle(evaluator, name):
    """Atoms that see different detunings: evaluate5_general_kernel (81-variable rho_11, vectors in shared-memory
    columns) against the dense 625x625 expm at the same per-atom offsets, in one batch with identical-atom rows
    (folded kernel) -- the planner splits the batch by kind.  Round 1 returned NaN rows here."""
    case = next(c for c in CASES if c["name"] == name)
    pts, seg, ns = C.pack_batch5([case] * 6)
    pts[1, 2], pts[1, 3] = 2 * np.pi * 41e3, -2 * np.pi * 23e3            # PT_OFF_A, PT_OFF_B
    pts[4, 2], pts[4, 3] = -2 * np.pi * 7e3, 2 * np.pi * 90e3
    out, st = evaluator.evaluate_host(5, case["noisy"], pts, seg, ns, want_states=True)
    assert np.isfinite(out[:, :14]).all()
    assert np.array_equal(out[0, :14], out[2, :14]) and np.array_equal(out[0, :14], out[5, :14])
    C.check_five_level_case(case, out[0], st[0])
    for i in (1, 4):
        fid, avg, info, fin = _oracle5(case, pts[i])
        assert abs(out[i, C.OUT["avg"]] - avg) < 1e-9
        assert abs(out[i, C.OUT["f01"]] - fid["01"]) < 1e-9 and abs(out[i, C.OUT["f10"]] - fid["10"]) < 1e-9
        assert abs(out[i, C.OUT["f11"]] - fid["11"]) < 1e-9
        ops = C.states_to_operators(5, st[i])
        if case["noisy"]:
            for key, lab, idx in (("rho11", "11", C.ACT2), ("rho01", "01", C.ACT01), ("rho10", "10", C.ACT10)):
                assert np.abs(ops[key][np.ix_(idx, idx)] - fin[lab][np.ix_(idx, idx)]).max() < 1e-10, key
        else:
            assert np.abs(ops["coh11"][C.ACT2, 0] - fin["11"][C.ACT2]).max() < 1e-10
    st_kind = evaluator.point_status(6)
    assert st_kind[1] == st_kind[4] and st_kind[0] != st_kind[1]


def test_shots_dim5_match_oracle_sampler_and_reduce(evaluator):
    """Monte-Carlo shots of the 5-level model: one perturbed row per shot (shot_rows_kernel, the 3-level model's
    Philox draws), rows with Doppler offsets on the general kernel; every shot equals the oracle at the
    oracle-sampled parameters, the device mean is the mean of the shots, sigma = 0 is the deterministic path."""
    from oracle import philox
    case = next(c for c in CASES if c["name"] == "lp5_rb_noisy")
    P, G, ns = C.pack_batch5([case] * 3)
    P[0, 9], P[0, 10] = 0.0, 0.0                                          # PT_SIG_DET, PT_SIG_RELR
    P[1, 9], P[1, 10] = 2 * np.pi * 30e3, 0.01
    P[2, 9], P[2, 10] = 2 * np.pi * 80e3, 0.03
    n_shots, seed = 4, 4321
    mean, shots = evaluator.evaluate_shots_host(5, True, n_shots, seed, P, G, ns, want_shots=True)
    det, _ = evaluator.evaluate_host(5, True, P, G, ns)
    assert np.abs(shots[0, :, :14] - det[0, :14]).max() < 1e-13
    for i in (1, 2):
        for s in range(n_shots):
            oa, ob, Vs = philox.shot_parameters(seed, i, s, P[i, 0], P[i, 9], P[i, 10])
            fid, avg, info, fin = _oracle5(case, P[i], V=Vs, offs=(oa, ob))
            assert abs(shots[i, s, C.OUT["avg"]] - avg) < 1e-9
            assert abs(shots[i, s, C.OUT["f11"]] - fid["11"]) < 1e-9
    assert np.abs(mean[:, :14] - shots[:, :, :14].mean(1)).max() < 1e-14


def test_public_api_rb_vs_cs(evaluator):
    """Config 5 through simulate_CZ_gate: Rb87 vs Cs133 with the intermediate state kept."""
    from noisyquantumsimulator_b200 import rydberg_gates as rg
    exc = rg.TwoPhotonExcitationConfig(laser_1=rg.LaserParameters(power=50e-6, waist=50e-6),
                                       laser_2=rg.LaserParameters(power=0.3, waist=50e-6), Delta_e=2 * np.pi * 1e9)
    inp = rg.LPSimulationInputs(excitation=exc, noise=rg.NoiseSourceConfig())
    for species, name in (("Rb87", "lp5_rb_noisy"), ("Cs133", "lp5_cs_noisy")):
        res = rg.simulate_CZ_gate(inp, species=species, hilbert_space_dim=5, tweezer_power=0.02, tweezer_waist=0.8e-6)
        exp = next(c for c in CASES if c["name"] == name)["expected"]
        assert abs(res.avg_fidelity - exp["avg_fidelity"]) < 1e-9
        assert res.results["11"].shape == (25, 25)


@pytest.mark.parametrize("proto", ["lp", "jp"])
def test_batch_device_prologue_equals_host_prologue_dim5(proto):
    """The pipelined column entry with prologue_kernel writing 20-double rows against the NumPy prologue +
    rydcz_evaluate_host, for the 5-level model."""
    from noisyquantumsimulator_b200 import rydberg_gates as rg
    from noisyquantumsimulator_b200.engine import get_evaluator
    rng = np.random.default_rng(23)
    B = 3000
    kw = dict(laser_1_power=50e-6, laser_1_waist=50e-6, laser_2_power=10 ** rng.uniform(-1.0, 0.3, B),
              laser_2_waist=50e-6, Delta_e=2 * np.pi * 1e9, tweezer_power=0.02, tweezer_waist=0.8e-6,
              temperature=rng.uniform(1e-6, 3e-5, B), spacing_factor=rng.uniform(2.5, 5.0, B), warn=False,
              hilbert_space_dim=5)
    if proto == "jp":
        kw.update(omega_tau=rng.uniform(5, 20, B), switching_times=np.sort(rng.uniform(1, 5, (B, 4)), 1),
                  phases=rng.uniform(-3, 3, (B, 5)))
    get_evaluator(0).set_chunk(1024)
    a = rg.simulate_CZ_gate_batch(proto, B=B, prologue="device", return_states=True, **kw)
    get_evaluator(0).set_chunk(65536)
    b = rg.simulate_CZ_gate_batch(proto, B=B, prologue="host", return_states=True, **kw)
    assert a.states.shape == (B, 129)
    assert np.abs(a.raw[:, :14] - b.raw[:, :14]).max() < 1e-9 and np.abs(a.states - b.states).max() < 1e-9
    assert np.allclose(a.tau_total, b.tau_total, rtol=1e-13)
