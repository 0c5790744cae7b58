"""5-level model (|0>, |1>, |e>, |r>, |l| per atom; 25x25 rho; BASELINE.json configs[4]) -- CPU side.

The reference has no 5-level code: the model is specified in DESIGN.md section 2.6, restated densely
in oracle/rydcz_oracle.py (hamiltonian5 / collapse_operators5 / evaluate5) and pinned only to that
oracle ("parity unpinned" against the reference) plus its own limit: for Delta_e -> infinity at
fixed effective Rabi frequency it must reproduce the reference's 3-level model."""
import numpy as np
import pytest

import _cases as C
from oracle import rydcz_oracle as orc

CASES = C.load("five_level_cases.json")


@pytest.fixture(scope="module")
def hostsim():
    return C.load_hostsim()


def test_oracle_reduces_to_three_level_model():
    """Adiabatic elimination: with balanced legs Omega_1 = |Omega_2| = sqrt(2 Delta_e Omega) the
    5-level result approaches the 3-level one like 1/Delta_e, noiseless and with Rydberg noise."""
    Om = 2 * np.pi * 3.6e6
    V, Delta, tau, d1 = 342.5 * Om, 0.377 * Om, 4.29 / Om, 2 * np.pi * 2e5
    segs = orc.segments_lp(Om, Delta, tau)
    r3 = orc.NoiseRates(gamma_r=7e3, gamma_phi_laser=1e5, gamma_loss_background=1e3)
    r5 = orc.NoiseRates5(g_r1=3.5e3, g_r0=3.5e3, g_rl=1e3, g_phi=1e5)
    for rates3, rates5 in ((None, None), (r3, r5)):
        _, a3, i3, _ = orc.evaluate(3, segs, V, d1, rates3)
        errs = []
        for De in (2 * np.pi * 1e10, 2 * np.pi * 1e11):
            _, a5, i5, _ = orc.evaluate5(segs, V, d1, np.sqrt(2 * De * Om), De, rates5)
            errs.append(abs(a5 - a3))
        assert errs[0] < 2e-3 and errs[1] < 1e-4 and errs[1] < errs[0] / 8


def test_reduced_generator_equals_dense_lindbladian():
    """codegen's reduced real systems reproduce the dense 625x625 Liouvillian on their subspace."""
    from noisyquantumsimulator_b200.codegen import model
    rng = np.random.default_rng(3)
    S = model.standard_systems(5)
    names = model.params_of(5)
    p = dict(om_re=rng.normal(), om_im=rng.normal(), det_a=rng.normal(), det_b=rng.normal(), V=3.0, delta1=0.2,
             g_r1=0.1, g_r0=0.07, g_phi=0.3, om1=1.7, det_e=4.0, g_e1=0.5, g_e0=0.4, g_rl=0.05)
    _, L = model.liouvillian_terms(5)
    dense = sum(p[n] * Lk for n, Lk in zip(names, L))
    for key, sysm in S.items():
        G = sysm.generator([p[n] for n in sysm.params])
        x = rng.normal(size=sysm.n)
        X = sysm.to_operator(x)
        dX = (dense @ X.reshape(-1)).reshape(25, 25)
        dx = G @ x
        got = sysm.to_operator(dx)
        mask = np.abs(sysm.to_operator(np.ones(sysm.n))) > 0        # compare on the kept entries only
        assert np.abs((got - dX)[mask]).max() < 1e-12, key
    sizes = {k: v.n for k, v in S.items()}
    assert sizes == {"rho01": 9, "coh01": 6, "rho10": 9, "coh10": 6, "rho11": 81, "coh11": 18}
    folded = model.symmetric_systems(5)
    assert folded["rho11s"].n == 45 and folded["coh11s"].n == 12


@pytest.mark.parametrize("noisy", [True, False])
def test_kernel_arithmetic_matches_oracle_fixtures(hostsim, noisy):
    """csrc/rydcz_core.cuh (host build): run_system_big with x and c in strided columns, the folded
    45-variable rho_11 system, the dim-5 parameter map, against the oracle's dense expm."""
    cases = [c for c in CASES if c["noisy"] == noisy]
    pts, seg, ns = C.pack_batch5(cases)
    out, st = hostsim(5, noisy, pts, seg, ns)
    for i, c in enumerate(cases):
        C.check_five_level_case(c, out[i], st[i], 1e-11, 1e-10)
        assert out[i, C.OUT["steps"]] > 0 and out[i, C.OUT["flops"]] > 0


def test_fixture_is_what_the_oracle_gives_today():
    """The committed fixture against a fresh oracle evaluation of one cheap case."""
    c = next(c for c in CASES if c["name"] == "lp5_rb_noisy")
    pt = np.array(c["point"])
    segs = [orc.Segment(s[0], complex(s[1], s[2]), s[3]) for s in c["segs"]]
    rates = orc.NoiseRates5(g_e1=pt[14], g_e0=pt[15], g_r1=pt[4], g_r0=pt[5], g_rl=pt[16], g_phi=pt[6])
    fid, avg, info, fin = orc.evaluate5(segs, pt[0], pt[1], pt[12], pt[13], rates, ls_comp=pt[17])
    assert abs(avg - c["expected"]["avg_fidelity"]) < 1e-12
    assert abs(np.trace(fin["11"]).real - 1.0) < 1e-12             # the full 25x25 rho keeps its trace


def test_public_api_dim5_on_host_harness(monkeypatch):
    """simulate_CZ_gate(hilbert_space_dim=5): same call, 25x25 results, Rb vs Cs (config 5)."""
    from noisyquantumsimulator_b200 import engine
    from noisyquantumsimulator_b200 import rydberg_gates as rg
    monkeypatch.setitem(engine._default, 0, C.HostsimEvaluator())
    exc = rg.TwoPhotonExcitationConfig(laser_1=rg.LaserParameters(power=50e-6, waist=50e-6),
                                       laser_2=rg.LaserParameters(power=0.3, waist=50e-6), Delta_e=2 * np.pi * 1e9)
    inp = rg.LPSimulationInputs(excitation=exc, noise=rg.NoiseSourceConfig())
    kw = dict(tweezer_power=0.02, tweezer_waist=0.8e-6)
    rb = rg.simulate_CZ_gate(inp, hilbert_space_dim=5, **kw)
    exp = next(c for c in CASES if c["name"] == "lp5_rb_noisy")["expected"]
    assert abs(rb.avg_fidelity - exp["avg_fidelity"]) < 1e-10
    assert rb.results["11"].shape == (25, 25) and len(rb.c_ops) == rb.noise_breakdown["n_collapse_ops"]
    assert rb.H1.shape == (25, 25) and np.allclose(rb.H1, rb.H1.conj().T)
    assert 0 < rb.phase_info["loss_probability"] < 0.2
    cs = rg.simulate_CZ_gate(inp, species="Cs133", hilbert_space_dim=5, **kw)
    exp = next(c for c in CASES if c["name"] == "lp5_cs_noisy")["expected"]
    assert abs(cs.avg_fidelity - exp["avg_fidelity"]) < 1e-10
    pure = rg.simulate_CZ_gate(inp, hilbert_space_dim=5, include_noise=False, **kw)
    assert pure.results["11"].shape == (25,) and abs(np.linalg.norm(pure.results["11"]) - 1) < 1e-11
    # batch API (device prologue by default)
    b = rg.simulate_CZ_gate_batch("lp", B=3, hilbert_space_dim=5, laser_1_power=50e-6, laser_1_waist=50e-6,
                                  laser_2_power=np.array([0.1, 0.3, 1.0]), laser_2_waist=50e-6,
                                  Delta_e=2 * np.pi * 1e9, warn=False, **kw)
    assert abs(b.avg_fidelity[1] - rb.avg_fidelity) < 1e-12 and b.raw.shape == (3, 16)



def _oracle5(case, pt):
    segs = [orc.Segment(s[0], complex(s[1], s[2]), s[3]) for s in case["segs"]]
    rates = orc.NoiseRates5(g_e1=pt[14], g_e0=pt[15], g_r1=pt[4], g_r0=pt[5], g_rl=pt[16], g_phi=pt[6]) \
        if case["noisy"] else None
    return orc.evaluate5(segs, pt[0], pt[1], pt[12], pt[13], rates, ls_comp=pt[17], detuning_offsets=(pt[2], pt[3]))


@pytest.mark.parametrize("name", ["lp5_rb_noisy", "lp5_rb_noiseless"])
def test_unequal_atoms_take_the_general_five_level_systems(hostsim, name):
    """Atoms that see different detunings (a Monte-Carlo shot's Doppler offsets): the general 81-variable rho_11 with
    its three Clenshaw vectors in columns (run_system_huge), the 18-variable coherence and both single-atom inputs,
    against the oracle's dense 625x625 expm at the same per-atom offsets.  (Round 1 returned NaN rows here.)"""
    case = next(c for c in CASES if c["name"] == name)
    pts, seg, ns = C.pack_batch5([case])
    pts[0, 2], pts[0, 3] = 2 * np.pi * 41e3, -2 * np.pi * 23e3            # PT_OFF_A, PT_OFF_B
    out, st = hostsim(5, case["noisy"], pts, seg, ns)
    fid, avg, info, fin = _oracle5(case, pts[0])
    assert abs(out[0, C.OUT["avg"]] - avg) < 1e-9
    assert abs(out[0, C.OUT["f01"]] - fid["01"]) < 1e-9 and abs(out[0, C.OUT["f10"]] - fid["10"]) < 1e-9
    assert abs(fid["01"] - fid["10"]) > 1e-7                              # the two atoms really differ
    ops = C.states_to_operators(5, st[0])
    if case["noisy"]:
        for key, lab, idx in (("rho11", "11", C.ACT2), ("rho01", "01", C.ACT01), ("rho10", "10", C.ACT10)):
            assert np.abs(ops[key][np.ix_(idx, idx)] - fin[lab][np.ix_(idx, idx)]).max() < 1e-10, key
        assert abs(out[0, C.OUT["loss"]] - (1.0 - np.trace(fin["11"][np.ix_(C.ACT2, C.ACT2)]).real)) < 1e-9
    else:
        assert np.abs(ops["coh11"][C.ACT2, 0] - fin["11"][C.ACT2]).max() < 1e-10
        assert np.abs(ops["coh10"][C.ACT10, 0] - fin["10"][C.ACT10]).max() < 1e-10
    # identical atoms through the general systems = the folded path
    pts[0, 2] = pts[0, 3] = 0.0
    a, sa = hostsim(5, case["noisy"], pts, seg, ns)
    b, sb = hostsim(5, case["noisy"], pts, seg, ns, allow_sym=False)
    assert np.abs(a[0, :14] - b[0, :14]).max() < 1e-11
