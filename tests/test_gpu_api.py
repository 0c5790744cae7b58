"""The reference-facing Python API on the real CUDA path (no evaluator substitution)."""
import numpy as np
import pytest

import _cases as C
from test_api_hostsim import check_result_against_golden

pytestmark = pytest.mark.gpu
REF = {c["name"]: c for c in C.load("reference_cases.json")}


@pytest.mark.parametrize("name", ["lp_noisy", "bb5_noisy", "sjp_noisy", "lp_noisy_cs133", "vb_bb5_default",
                                  "vb_sjp_default", "lp_dim4_noisy", "lp_noisy_neg_Delta_e", "lp_shaped_cosine_noisy"])
@pytest.mark.parametrize("phase_mode", ["coherence", "compat"])
def test_simulate_cz_gate(name, phase_mode):
    from noisyquantumsimulator_b200 import rydberg_gates as rg
    case = REF[name]
    inputs, kw = C.api_inputs(case)
    res = rg.simulate_CZ_gate(inputs, phase_mode=phase_mode, **kw)
    check_result_against_golden(res, case, phase_mode)


def test_known_answers_through_public_api():
    """The notebook's saved QuTiP numbers via simulate_CZ_gate itself."""
    from noisyquantumsimulator_b200 import rydberg_gates as rg
    for case in C.load("known_answers.json"):
        inputs, kw = C.api_inputs(case)
        res = rg.simulate_CZ_gate(inputs, **kw)
        exp, tol = case["expected_printed"], case["tolerance_printed"]
        got = dict(avg_fidelity=res.avg_fidelity, f11=res.fidelities["11"],
                   cz_phase_fidelity=res.phase_info["cz_phase_fidelity"],
                   controlled_phase_deg=res.phase_info["controlled_phase_deg"],
                   phase_error_deg=res.phase_info["phase_error_from_pi_deg"], gate_time_us=res.gate_time_us,
                   V_over_Omega=res.V_over_Omega, Omega_MHz=res.Omega_MHz)
        for k, v in exp.items():
            if k != "cost":
                assert abs(got[k] - v) <= tol[k], (case["name"], k)


def test_sweep_batch_with_shots_converges_to_deterministic_limit():
    """Config 3 shape in miniature: a power x detuning grid with Monte-Carlo shots.  At T -> 0 the
    shot average equals the deterministic evaluation; at finite T the shot mean of F is below it
    and its scatter shrinks like 1/sqrt(N)."""
    from noisyquantumsimulator_b200 import rydberg_gates as rg
    p2, doo = np.meshgrid(np.geomspace(0.05, 2.0, 6), np.linspace(0.2, 0.5, 6), indexing="ij")
    base = dict(laser_1_power=50e-6, laser_1_waist=50e-6, laser_2_power=p2.ravel(), laser_2_waist=50e-6,
                Delta_e=2 * np.pi * 1e9, tweezer_power=0.02, tweezer_waist=0.8e-6,
                delta_over_omega=doo.ravel(), omega_tau=4.3)
    cold = rg.simulate_CZ_gate_batch("lp", B=36, temperature=1e-16, n_shots=16, seed=1234, **base)
    det = rg.simulate_CZ_gate_batch("lp", B=36, temperature=1e-16, include_doppler_dephasing=False,
                                    include_motional_dephasing=False, **base)
    # (the shot path runs the general kernel, the deterministic one the exchange-folded kernel)
    assert np.abs(cold.avg_fidelity - det.avg_fidelity).max() < 1e-8
    hot_a = rg.simulate_CZ_gate_batch("lp", B=36, temperature=20e-6, n_shots=64, seed=1, **base)
    hot_b = rg.simulate_CZ_gate_batch("lp", B=36, temperature=20e-6, n_shots=64, seed=2, **base)
    hot_c = rg.simulate_CZ_gate_batch("lp", B=36, temperature=20e-6, n_shots=1024, seed=3, **base)
    hot_d = rg.simulate_CZ_gate_batch("lp", B=36, temperature=20e-6, n_shots=1024, seed=4, **base)
    assert np.all(hot_c.fidelities[:, 1] <= 1.0)
    s64 = np.abs(hot_a.avg_fidelity - hot_b.avg_fidelity).mean()
    s1024 = np.abs(hot_c.avg_fidelity - hot_d.avg_fidelity).mean()
    assert s1024 < s64 / 2                      # ~1/4 expected
    same = rg.simulate_CZ_gate_batch("lp", B=36, temperature=20e-6, n_shots=64, seed=1, **base)
    assert np.array_equal(same.avg_fidelity, hot_a.avg_fidelity)   # counter-based: reproducible
