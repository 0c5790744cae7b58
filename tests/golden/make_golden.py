#!/usr/bin/env python3
"""Generate tests/golden/*.json by running the REFERENCE's own simulate_CZ_gate.

Runs only in the build container (needs /root/reference).  QuTiP is replaced by
``qutip_shim`` (NumPy + the oracle's restatement of mesolve); every other line
executed is the reference's unmodified code.  Outputs:

  known_answers.json   -- the 13 saved QuTiP outputs of the reference's notebook
                          (RG/cz_gate_optimization_demo.ipynb cells 6, 10, 18, 22),
                          transcribed by hand, with the inputs that produce them.
                          make_golden asserts the shim run reproduces each one.
  reference_cases.json -- ~45 cases (LP / bang-bang / smooth-JP / shaped, noisy and
                          noiseless, Rb/Cs, 3- and 4-level, quirks) with the
                          reference's scalars, segment schedule, rates, fidelity
                          outputs and final states.

Usage:  python tests/golden/make_golden.py
"""
from __future__ import annotations

import json
import math
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import qutip_shim as qs  # noqa: E402
from oracle import rydcz_oracle as orc  # noqa: E402

rg = qs.import_reference()
sim, cfg = rg.simulation, rg.configurations

# -- record every mesolve call so the segment schedule can be cross-checked ----
_calls = []
_orig_mesolve = qs.mesolve


def _recording_mesolve(H, psi0, tlist, c_ops=None, options=None, **kw):
    _calls.append((H.data.copy(), float(tlist[-1]), len(tlist), len(c_ops or []),
                   dict(options or {})))
    return _orig_mesolve(H, psi0, tlist, c_ops=c_ops, options=options, **kw)


sim.mesolve = _recording_mesolve

MEDIUM = dict(species="Rb87", n_rydberg=70, temperature=2e-6, spacing_factor=2.8,
              tweezer_power=0.020, tweezer_waist=0.8e-6, B_field=1e-4, NA=0.5)
NOISELESS = dict(include_spontaneous_emission=False, include_intermediate_scattering=False,
                 include_motional_dephasing=False, include_doppler_dephasing=False,
                 include_intensity_noise=False, intensity_noise_frac=0.0,
                 include_laser_dephasing=False, include_magnetic_dephasing=False)
FULL = dict(include_spontaneous_emission=True, include_intermediate_scattering=True,
            include_motional_dephasing=True, include_doppler_dephasing=True,
            include_intensity_noise=True, intensity_noise_frac=0.01,
            include_laser_dephasing=True, include_magnetic_dephasing=True)


def exc_dict(p2=0.3, p1=50e-6, w1=50e-6, w2=50e-6, lw=100.0, pur=1.0, Delta_e=2 * np.pi * 1e9,
             counter=True):
    return dict(laser_1=dict(power=p1, waist=w1, polarization="pi", polarization_purity=pur,
                             linewidth_hz=lw),
                laser_2=dict(power=p2, waist=w2, polarization="sigma+", polarization_purity=pur,
                             linewidth_hz=lw),
                Delta_e=Delta_e, counter_propagating=counter)


def build_inputs(api):
    e = api["excitation"]
    excitation = cfg.TwoPhotonExcitationConfig(
        laser_1=cfg.LaserParameters(**e["laser_1"]), laser_2=cfg.LaserParameters(**e["laser_2"]),
        Delta_e=e["Delta_e"], counter_propagating=e["counter_propagating"])
    noise = cfg.NoiseSourceConfig(**api["noise"])
    cls = {"lp": cfg.LPSimulationInputs, "jp": cfg.JPSimulationInputs,
           "sjp": cfg.SmoothJPSimulationInputs}[api["protocol"]]
    return cls(excitation=excitation, noise=noise, **api["params"])


def numeric(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, (bool, np.bool_)):
            out[k] = bool(v)
        elif isinstance(v, (int, float, np.integer, np.floating)):
            out[k] = float(v)
        elif isinstance(v, str):
            out[k] = v
        elif isinstance(v, (list, tuple)) and all(isinstance(x, (int, float)) for x in v):
            out[k] = [float(x) for x in v]
    return out


def cplx(a):
    a = np.asarray(a)
    return [np.real(a).ravel().tolist(), np.imag(a).ravel().tolist()]


def run_case(name, protocol, params, kwargs, excitation=None, noise=None, include_noise=False,
             store_states=True):
    api = dict(protocol=protocol, params=params, excitation=excitation or exc_dict(),
               noise=noise or (FULL if include_noise else NOISELESS),
               kwargs=dict(kwargs, include_noise=include_noise))
    inputs = build_inputs(api)
    _calls.clear()
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        res = sim.simulate_CZ_gate(simulation_inputs=inputs, verbose=False, return_dataclass=True,
                                   **api["kwargs"])
    dim = res.hilbert_space_dim
    # ---- reconstruct the schedule the reference used from its recorded mesolve calls
    n_seg = len(_calls) // 4
    assert len(_calls) == 4 * n_seg
    calls = _calls[:n_seg]
    nb = res.noise_breakdown
    rates = None
    if include_noise:
        rates = orc.NoiseRates(
            gamma_r=nb["gamma_r"], gamma_bbr=nb["gamma_bbr"], gamma_phi_laser=nb["gamma_phi_laser"],
            gamma_phi_thermal=nb["gamma_phi_thermal"], gamma_phi_zeeman=nb["gamma_phi_zeeman"],
            gamma_loss_antitrap=nb["gamma_loss_antitrap"],
            gamma_loss_background=nb["gamma_loss_background"],
            gamma_scatter_intermediate=nb["gamma_scatter_intermediate"],
            gamma_leakage=nb["gamma_leakage"], mJ_leakage_rate=nb["mJ_leakage_rate"],
            branching_1=nb["branching_1"])
        assert len(orc.collapse_operators(dim, rates)) == nb["n_collapse_ops"] == calls[0][3]
    delta_1 = res.delta_zeeman + (res.delta_stark if res.trap_laser_on else 0.0)
    segs = []
    r_idx = 2  # |r> (or |r+>) level index on one atom
    for Hm, t_end, n_t, n_c, opts in calls:
        # read Ω, Δ off the reference's own matrix: <0 r|H|0 1> = Ω/2, <0 r|H|0 r> = -Δ
        Om = 2.0 * Hm[r_idx, 1]
        De = -np.real(Hm[r_idx, r_idx])
        seg = orc.Segment(t_end, complex(Om), float(De), n_t)
        Hchk = orc.hamiltonian(dim, seg.Omega, seg.Delta, res.V, delta_1)
        assert np.allclose(Hchk, Hm, rtol=1e-13, atol=1e-6), name
        segs.append(seg)
    # ---- the oracle, driven standalone, must agree with the shim-driven reference run
    fid, avg, info, finals = orc.evaluate(dim, segs, res.V, delta_1, rates, "exact", "compat")
    assert abs(avg - res.avg_fidelity) < 1e-11, (name, avg, res.avg_fidelity)
    rec = dict(
        name=name, api=api, dim=dim,
        warnings=[str(w.message)[:60] for w in wlist if issubclass(w.category, UserWarning)],
        scalars=numeric(dict(
            Omega=res.Omega, V=res.V, Delta=res.Delta, V_over_Omega=res.V_over_Omega,
            Delta_over_Omega=res.Delta_over_Omega, tau_single=res.tau_single,
            tau_total=res.tau_total, R=res.R, xi_re=np.real(res.xi), xi_im=np.imag(res.xi),
            U0_mK=res.U0_mK, omega_r_kHz=res.omega_r_kHz, sigma_r_nm=res.sigma_r_nm,
            trap_wavelength_nm=res.trap_wavelength_nm, delta_zeeman=res.delta_zeeman,
            delta_stark=res.delta_stark, n_pulses=res.n_pulses, delta_1=delta_1)),
        protocol_name=res.protocol,
        noise_breakdown=numeric(nb), magic=numeric(res.magic_wavelength_analysis),
        pulse_info=numeric(res.pulse_info),
        segments=dict(dt=[s.dt for s in segs], re=[s.Omega.real for s in segs],
                      im=[s.Omega.imag for s in segs], delta=[s.Delta for s in segs],
                      n_tlist=[s.n_tlist for s in segs]),
        rates=None if rates is None else rates.__dict__,
        result=dict(avg_fidelity=float(res.avg_fidelity),
                    fidelities={k: float(v) for k, v in res.fidelities.items()},
                    phase_info=numeric(res.phase_info)),
    )
    # gauge-free epilogue (this repo's default for the noisy branch), from the oracle
    if include_noise:
        f2, a2, i2, _ = orc.evaluate(dim, segs, res.V, delta_1, rates, "exact", "coherence")
        rec["result_coherence"] = dict(avg_fidelity=a2, fidelities=f2, phase_info=numeric(i2))
    if store_states:
        rec["finals"] = {k: cplx(res.results[k].full()) for k in orc.LABELS}
    return rec, res


def main():
    # ------------------------------------------------------------------ known answers
    ka_specs = [
        # name, protocol, params, kwargs overrides, excitation, expected
        ("lp_baseline", "lp", {}, {}, exc_dict(),
         dict(avg_fidelity=0.994423, f11=0.977897, cz_phase_fidelity=0.978587,
              controlled_phase_deg=163.17, phase_error_deg=16.83, gate_time_us=0.379,
              V_over_Omega=342.5, Omega_MHz=3.602, cost=36.7118)),
    ]
    for mw, F, vo, t, err in [(50, 0.969304, 839.0, 0.929, 40.21), (100, 0.983722, 593.2, 0.657, 29.01),
                              (200, 0.991645, 419.5, 0.465, 20.66), (300, 0.994423, 342.5, 0.379, 16.83),
                              (500, 0.996704, 265.3, 0.294, 12.90), (1000, 0.998474, 187.6, 0.208, 8.72),
                              (2000, 0.999423, 132.6, 0.147, 5.23)]:
        ka_specs.append((f"lp_power_{mw}mW", "lp", {}, {}, exc_dict(p2=mw * 1e-3),
                         dict(avg_fidelity=F, V_over_Omega=vo, gate_time_us=t, phase_error_deg=err)))
    ka_specs += [
        ("lp_optimum", "lp", dict(delta_over_omega=0.369547, omega_tau=4.215360),
         dict(spacing_factor=5.373584), exc_dict(),
         dict(avg_fidelity=0.999918, f11=0.999995, cz_phase_fidelity=0.999998,
              controlled_phase_deg=179.83, gate_time_us=0.372, V_over_Omega=6.9)),
        ("sjp_baseline", "sjp", {}, {}, exc_dict(),
         dict(avg_fidelity=0.992310, f11=0.969240, cz_phase_fidelity=0.969287,
              controlled_phase_deg=159.81, gate_time_us=0.446, V_over_Omega=342.5, cost=72.0941)),
        ("sjp_optimum", "sjp", dict(omega_tau=9.689759, A=0.964291, omega_mod_ratio=1.297048,
                                    phi_offset=1.555215, delta_over_omega=0.064792),
         dict(spacing_factor=2.386530), exc_dict(),
         dict(avg_fidelity=0.999996, f11=1.000000, controlled_phase_deg=179.96,
              gate_time_us=0.428, V_over_Omega=893.3)),
        ("sjp_two_phase_optimum", "sjp", dict(omega_tau=9.684057, A=0.977122,
                                              omega_mod_ratio=1.295789, phi_offset=1.705593,
                                              delta_over_omega=0.069904),
         {}, exc_dict(),
         dict(avg_fidelity=0.999544, f11=0.999815, cz_phase_fidelity=0.999898,
              controlled_phase_deg=-178.84, gate_time_us=0.428, V_over_Omega=342.5)),
    ]
    known = []
    for name, proto, params, kov, exc, expect in ka_specs:
        rec, res = run_case("ka_" + name, proto, params, dict(MEDIUM, **kov), excitation=exc,
                            store_states=False)
        got = dict(avg_fidelity=res.avg_fidelity, f11=res.fidelities["11"],
                   cz_phase_fidelity=res.phase_info["cz_phase_fidelity"],
                   controlled_phase_deg=res.phase_info["controlled_phase_deg"],
                   phase_error_deg=res.phase_info["phase_error_from_pi_deg"],
                   gate_time_us=res.gate_time_us, V_over_Omega=res.V_over_Omega,
                   Omega_MHz=res.Omega_MHz)
        got["cost"] = orc.compute_cost(got["avg_fidelity"], got["f11"], got["cz_phase_fidelity"],
                                       got["gate_time_us"])
        tol = dict(avg_fidelity=6e-7, f11=6e-7, cz_phase_fidelity=6e-7, controlled_phase_deg=6e-3,
                   phase_error_deg=6e-3, gate_time_us=6e-4, V_over_Omega=6e-2, Omega_MHz=6e-4,
                   cost=6e-5)
        for k, v in expect.items():
            assert abs(got[k] - v) <= tol[k], (name, k, got[k], v)
        rec["expected_printed"] = expect
        rec["tolerance_printed"] = {k: tol[k] for k in expect}
        known.append(rec)
        print(f"known-answer {name}: F={res.avg_fidelity:.6f} OK")
    json.dump(dict(source="RG/cz_gate_optimization_demo.ipynb cells 6,10,18,22 (saved QuTiP outputs)",
                   cases=known), open(os.path.join(HERE, "known_answers.json"), "w"), indent=0)

    # ------------------------------------------------------------------ reference cases
    cases = []

    def add(name, proto, params, kov=None, exc=None, noise=None, include_noise=False):
        rec, res = run_case(name, proto, params, dict(MEDIUM, **(kov or {})), excitation=exc,
                            noise=noise, include_noise=include_noise)
        cases.append(rec)
        print(f"case {name}: dim={rec['dim']} S={len(rec['segments']['dt'])} "
              f"F={res.avg_fidelity:.6f}")

    pi = math.pi
    noisy_exc = exc_dict(pur=0.99)
    BB5 = dict(omega_tau=22.08, switching_times=[2.214, 8.823, 13.258, 19.867],
               phases=[pi / 2, 0, -pi / 2, 0, pi / 2])
    BB7 = dict(omega_tau=7.0, switching_times=[0.3328, 0.5859, 3.4340, 3.5530, 4.1204, 6.7431],
               phases=[pi / 2, 0, -pi / 2, -pi / 2, 0, pi / 2, 0])
    VB = dict(tweezer_power=30e-3, tweezer_waist=1.0e-6)  # validate_bangbang.py uses the defaults
    vb_exc = exc_dict(lw=100.0, pur=0.99)  # LaserParameters defaults there
    # validate_bangbang.py:66-99 (no saved output in the reference: parity unpinned)
    add("vb_bb5_default", "jp", BB5, VB, vb_exc)
    add("vb_bb5_perturbed", "jp", dict(BB5, switching_times=[3.0, 9.0, 14.0, 20.0]), VB, vb_exc)
    add("vb_bb5_quarter_phases", "jp", dict(BB5, phases=[pi / 4, 0, -pi / 4, 0, pi / 4]), VB, vb_exc)
    add("vb_bb7_paper", "jp", BB7, VB, vb_exc)
    add("vb_sjp_default", "sjp", {}, VB, vb_exc)
    add("vb_bb_none_defaults", "jp", {}, VB, vb_exc)
    # noisy baselines, medium apparatus
    add("lp_noisy", "lp", {}, None, noisy_exc, include_noise=True)
    add("sjp_noisy", "sjp", {}, None, noisy_exc, include_noise=True)
    add("bb5_noisy", "jp", BB5, None, noisy_exc, include_noise=True)
    add("bb7_noisy", "jp", BB7, None, noisy_exc, include_noise=True)
    add("lp_noisy_override", "lp", dict(delta_over_omega=0.31, omega_tau=4.9), None, noisy_exc,
        include_noise=True)
    # environment / apparatus variations (noisy)
    add("lp_noisy_hot50uK", "lp", {}, dict(temperature=50e-6), noisy_exc, include_noise=True)
    add("lp_noisy_spacing2", "lp", {}, dict(spacing_factor=2.0), noisy_exc, include_noise=True)
    add("lp_noisy_spacing5p4_weak", "lp", {}, dict(spacing_factor=5.4), noisy_exc, include_noise=True)
    add("lp_noisy_spacing6p5_vlt10", "lp", {}, dict(spacing_factor=6.5), noisy_exc, include_noise=True)
    add("lp_noisy_n50", "lp", {}, dict(n_rydberg=50), noisy_exc, include_noise=True)
    add("lp_noisy_n90", "lp", {}, dict(n_rydberg=90), noisy_exc, include_noise=True)
    add("lp_noisy_cs133", "lp", {}, dict(species="Cs133", qubit_0=(3, 0), qubit_1=(4, 0)),
        noisy_exc, include_noise=True)
    add("sjp_noisy_cs133", "sjp", {}, dict(species="Cs133", qubit_0=(3, 0), qubit_1=(4, 0)),
        noisy_exc, include_noise=True)
    add("lp_noisy_linewidth_1MHz", "lp", {}, None, exc_dict(lw=1e6, pur=0.99), include_noise=True)
    add("lp_noisy_wavelength820", "lp", {}, dict(tweezer_wavelength_nm=820.0), noisy_exc,
        include_noise=True)
    add("lp_noisy_trap_off", "lp", {}, dict(trap_laser_on=False), noisy_exc, include_noise=True)
    add("lp_noisy_bg_loss", "lp", {}, dict(background_loss_rate_hz=2.5e4), noisy_exc,
        include_noise=True)
    add("lp_noisy_stretched_qubit", "lp", {}, dict(qubit_0=(1, 1), qubit_1=(2, 2)), noisy_exc,
        include_noise=True)
    add("lp_noisy_copropagating", "lp", {}, None, exc_dict(pur=0.99, counter=False),
        include_noise=True)
    add("lp_noisy_neg_Delta_e", "sjp", {}, None, exc_dict(pur=0.99, Delta_e=-2 * np.pi * 1e9),
        include_noise=True)
    add("lp_noisy_bigB", "lp", {}, dict(B_field=5e-4), noisy_exc, include_noise=True)
    add("lp_noisy_deep_trap", "lp", {}, dict(tweezer_power=0.1, tweezer_waist=0.6e-6), noisy_exc,
        include_noise=True)
    add("lp_noisy_NA07", "lp", {}, dict(NA=0.7), noisy_exc, include_noise=True)
    for tag, flags in [("no_motional", dict(include_motional_dephasing=False)),
                       ("no_doppler", dict(include_doppler_dephasing=False)),
                       ("no_intensity", dict(include_intensity_noise=False)),
                       ("intensity5pct", dict(intensity_noise_frac=0.05))]:
        add(f"lp_noisy_{tag}", "lp", {}, None, noisy_exc, noise=dict(FULL, **flags),
            include_noise=True)
    # include_noise=True but a 'noiseless' NoiseSourceConfig: most flags are dead
    add("lp_include_noise_flags_off", "lp", {}, None, noisy_exc, noise=NOISELESS, include_noise=True)
    # bang-bang edge cases
    add("bb5_overshoot_switch", "jp", dict(BB5, switching_times=[2.214, 8.823, 13.258, 23.0]),
        None, noisy_exc)
    add("bb5_zero_length_segment", "jp", dict(BB5, switching_times=[2.214, 8.823, 8.823, 19.867]),
        None, noisy_exc)
    add("bb5_random", "jp", dict(omega_tau=11.3, switching_times=[1.1, 2.9, 6.4, 9.9],
                                 phases=[0.3, -2.1, 1.7, 3.0, -0.4]), None, noisy_exc,
        include_noise=True)
    # smooth-JP quirks: A=0.0 / phi_offset=0.0 / ratio=0.0 fall back to defaults (getattr ... or ...)
    add("sjp_zero_params_fallback", "sjp", dict(A=0.0, omega_mod_ratio=0.0, phi_offset=0.0,
                                                omega_tau=9.0, delta_over_omega=0.03),
        None, noisy_exc)
    add("sjp_custom_noisy", "sjp", dict(omega_tau=12.0, A=1.3, omega_mod_ratio=0.9, phi_offset=0.7,
                                        delta_over_omega=0.05), None, noisy_exc, include_noise=True)
    # shaped LP pulses (scalar-envelope quirk: gaussian/blackman degenerate to square amplitude)
    for shape in ("gaussian", "cosine", "blackman"):
        add(f"lp_shaped_{shape}", "lp", dict(pulse_shape=shape), None, noisy_exc)
    add("lp_shaped_cosine_noisy", "lp", dict(pulse_shape="cosine"), None, noisy_exc,
        include_noise=True)
    # 4-level model
    add("lp_dim4_noiseless", "lp", {}, dict(hilbert_space_dim=4), noisy_exc)
    add("lp_dim4_noisy", "lp", {}, dict(hilbert_space_dim=4), noisy_exc, include_noise=True)
    add("bb5_dim4_noisy", "jp", BB5, dict(hilbert_space_dim=4), exc_dict(pur=0.95),
        include_noise=True)
    json.dump(dict(generator="tests/golden/make_golden.py (reference orchestration + oracle numerics)",
                   cases=cases), open(os.path.join(HERE, "reference_cases.json"), "w"), indent=0)
    print(f"wrote {len(known)} known answers, {len(cases)} reference cases")


if __name__ == "__main__":
    main()
