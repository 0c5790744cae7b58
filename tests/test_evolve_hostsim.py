"""The reference's evolve_* helpers and operator builders (RG/simulation.py:647-776, 1502-2235;
RG/noise_models.py:1199-1446) served by the kernel arithmetic.  CPU: host harness; the same calls run on the
CUDA path in tests/test_gpu_api.py."""
import numpy as np
import pytest

import _cases as C

REF = {c["name"]: c for c in C.load("reference_cases.json")}


@pytest.fixture()
def rg(monkeypatch):
    from noisyquantumsimulator_b200 import engine
    from noisyquantumsimulator_b200 import rydberg_gates
    fake = C.HostsimEvaluator()
    monkeypatch.setattr(engine, "get_evaluator", lambda device=0: fake)
    return rydberg_gates


def _kets(rg, dim):
    hs = rg.build_hilbert_space(dim)
    one = {"0": hs.basis["0"], "1": hs.basis["1"]}
    return hs, {a + b: np.kron(one[a], one[b]) for a in "01" for b in "01"}


def _c_ops(rg, case, hs):
    if case["rates"] is None:
        return []
    nb = case["noise_breakdown"]
    c_ops, _ = rg.build_all_noise_operators(
        hs=hs, gamma_r=nb["gamma_r"], gamma_bbr=nb["gamma_bbr"], gamma_phi_laser=nb["gamma_phi_laser"],
        gamma_phi_thermal=nb["gamma_phi_thermal"], gamma_phi_zeeman=nb["gamma_phi_zeeman"],
        gamma_loss_antitrap=nb["gamma_loss_antitrap"], gamma_loss_background=nb["gamma_loss_background"],
        gamma_scatter_intermediate=nb["gamma_scatter_intermediate"], gamma_leakage=nb["gamma_leakage"],
        mJ_leakage_rate=nb["mJ_leakage_rate"])
    return c_ops


def _check(final, case):
    fin = C.golden_finals(case)
    for lab in ("00", "01", "10", "11"):
        got = np.asarray(final[lab])
        assert np.abs(got.reshape(fin[lab].shape) - fin[lab]).max() < 1e-10, lab


@pytest.mark.parametrize("name", ["lp_noisy", "lp_dim4_noiseless", "lp_dim4_noisy"])
def test_evolve_two_pulse_lp_reproduces_the_reference_run(rg, name):
    from noisyquantumsimulator_b200.rydberg_gates import simulation as sim
    case = REF[name]
    hs, kets = _kets(rg, case["dim"])
    sc = case["scalars"]
    xi = complex(sc["xi_re"], sc["xi_im"])
    hk = dict(Delta=sc["Delta"], V=sc["V"], hs=hs, delta_zeeman=sc["delta_zeeman"], delta_stark=sc["delta_stark"])
    H1 = rg.build_full_hamiltonian(Omega=sc["Omega"], **hk)
    H2 = rg.build_full_hamiltonian(Omega=sc["Omega"] * xi, **hk)
    c_ops = _c_ops(rg, case, hs)
    final = sim.evolve_two_pulse_lp(kets, H1, H2, sc["tau_single"], c_ops)
    _check(final, case)
    one = sim.evolve_two_pulse(H1, H2, kets["11"], sc["tau_single"], c_ops)
    assert np.abs(np.asarray(one).reshape(np.asarray(final["11"]).shape) - final["11"]).max() < 1e-13
    # a single mesolve call = a single exact segment
    mid = sim.evolve_state(H1, kets["01"], np.linspace(0, sc["tau_single"], 100), c_ops)
    from oracle import rydcz_oracle as orc
    rates = orc.NoiseRates(**case["rates"]) if case["rates"] else None
    fin = orc.propagate(case["dim"], [orc.Segment(sc["tau_single"], complex(sc["Omega"]), sc["Delta"])], sc["V"],
                        sc["delta_zeeman"] + sc["delta_stark"], rates, mode="exact")
    assert np.abs(np.asarray(mid).reshape(np.asarray(fin["01"]).shape) - fin["01"]).max() < 1e-10


@pytest.mark.parametrize("name", ["bb5_noisy", "bb5_overshoot_switch"])
def test_evolve_bangbang_jp_reproduces_the_reference_run(rg, name):
    from noisyquantumsimulator_b200.rydberg_gates import simulation as sim
    case = REF[name]
    hs, kets = _kets(rg, case["dim"])
    sc, api = case["scalars"], case["api"]["params"]
    final, info = sim.evolve_bangbang_jp(kets, sc["Omega"], sc["V"], api["omega_tau"], api["switching_times"],
                                         api["phases"], hs, _c_ops(rg, case, hs), delta_zeeman=sc["delta_zeeman"],
                                         delta_stark=sc["delta_stark"])
    _check(final, case)
    assert info["n_segments"] == len(api["phases"]) and info["protocol"] == "bangbang_jp"
    with pytest.raises(AssertionError, match="Need len"):
        sim.evolve_bangbang_jp(kets, sc["Omega"], sc["V"], 7.0, [1.0], [0.0, 0.1, 0.2], hs)


def test_evolve_smooth_jp_reproduces_the_reference_run(rg):
    from noisyquantumsimulator_b200.rydberg_gates import simulation as sim
    case = REF["sjp_custom_noisy"]
    hs, kets = _kets(rg, 3)
    sc, p = case["scalars"], case["api"]["params"]
    final, info = sim.evolve_smooth_sinusoidal_jp(
        kets, sc["Omega"], sc["V"], sc["tau_total"], hs, _c_ops(rg, case, hs), A=p["A"],
        omega_mod=p["omega_mod_ratio"] * sc["Omega"], phi_offset=p["phi_offset"],
        delta_over_omega=case["pulse_info"]["delta_over_omega"], n_steps=300, delta_zeeman=sc["delta_zeeman"],
        delta_stark=sc["delta_stark"])
    _check(final, case)
    assert info["n_steps"] == 300 and len(info["phase_evolution"]) == 300


def test_operators_outside_the_model_are_refused(rg):
    from noisyquantumsimulator_b200.rydberg_gates import simulation as sim
    from noisyquantumsimulator_b200.rydberg_gates import evolution as evo
    hs, kets = _kets(rg, 3)
    H = rg.build_full_hamiltonian(2e7, 1e6, 7e9, hs)
    bad = H.copy()
    bad[0, 4] = bad[4, 0] = 1e5                      # couples |00> and |11>: not a term of the model
    with pytest.raises(ValueError, match="outside the model"):
        sim.evolve_state(bad, kets["11"], [0, 1e-7])
    with pytest.raises(ValueError, match="outside the channel set"):
        sim.evolve_state(H, kets["11"], [0, 1e-7], [np.sqrt(1e3) * np.kron(hs.transitions["1->r"], hs.identity)])
    psi = (kets["01"] + kets["10"]) / np.sqrt(2)
    with pytest.raises(NotImplementedError, match="computational-basis inputs"):
        sim.evolve_state(H, psi, [0, 1e-7])
    p = evo.hamiltonian_parameters(H, 3)
    assert abs(p["V"] - 7e9) < 1e-5 and abs(p["om_re"] - 2e7) < 1e-7 and abs(p["det_a"] - 1e6) < 1e-4


def test_operator_builders_match_the_combined_list(rg):
    """build_decay / dephasing / loss / scatter operators (RG/noise_models.py:1199-1446) against
    build_all_noise_operators, which the golden fixtures pin to the reference's list."""
    from noisyquantumsimulator_b200.rydberg_gates import evolution as evo
    for dim in (3, 4):
        hs = rg.build_hilbert_space(dim)
        parts = (rg.build_decay_operators(3.0e3, hs, gamma_bbr=1.0e3, branching_1=0.4, leakage_rate=50.0)
                 + rg.build_dephasing_operators(2.0e4, hs) + rg.build_loss_operators(70.0, hs)
                 + rg.build_scatter_operators(9.0e3, hs))
        full, nb = rg.build_all_noise_operators(hs=hs, gamma_r=3.0e3, gamma_bbr=1.0e3, gamma_phi_laser=2.0e4,
                                                gamma_loss_antitrap=70.0, gamma_scatter_intermediate=9.0e3,
                                                branching_1=0.4, mJ_leakage_rate=50.0)
        assert len(parts) == len(full) == nb["n_collapse_ops"] == (12 if dim == 3 else 26)
        a, b = evo.collapse_rates(parts, dim), evo.collapse_rates(full, dim)
        assert all(abs(a[k] - b[k]) < 1e-9 * max(b[k], 1.0) for k in b)
        # dim 3, reference order: r->1 (A, B), r->0 (A, B), bbr (A, B), ...
        if dim == 3:
            assert np.allclose(parts[0], np.sqrt(3.0e3 * 0.4) * np.kron(hs.transitions["r->1"], hs.identity))
            assert np.allclose(parts[3], np.sqrt(3.0e3 * 0.6) * np.kron(hs.identity, hs.transitions["r->0"]))
    assert rg.build_dephasing_operators(0.0, hs) == [] and rg.build_loss_operators(-1.0, hs) == []
    q = rg.build_loss_operators(5.0, rg.build_hilbert_space(3), loss_source="qubit")
    assert len(q) == 2 and q[0][0, 3] == pytest.approx(np.sqrt(5.0))      # |00><10|
