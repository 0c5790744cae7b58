"""Golden fixtures for the 5-level model (hilbert_space_dim=5; BASELINE.json configs[4]).

The reference has no 5-level code, so these are ORACLE outputs (oracle/rydcz_oracle.py: dense
25x25 Lindblad, scipy expm) on inputs produced by the product's own scalar prologue
(rydberg_gates/prologue.build_plan, which is pinned to the reference's prologue for every scalar
it shares with the 3-level path).  Parity for this model is therefore UNPINNED against the
reference; the fixtures pin the CUDA path to the oracle and guard against regressions.

    python tests/golden/make_golden5.py        # rewrites tests/golden/five_level_cases.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from noisyquantumsimulator_b200.rydberg_gates import prologue  # noqa: E402
from oracle import rydcz_oracle as orc  # noqa: E402

MEDIUM = dict(laser_1_power=50e-6, laser_1_waist=50e-6, laser_2_power=0.3, laser_2_waist=50e-6,
              Delta_e=2 * np.pi * 1e9, temperature=2e-6, spacing_factor=2.8, n_rydberg=70, species="Rb87",
              tweezer_power=0.020, tweezer_waist=0.8e-6, B_field=1e-4, NA=0.5)

CASES = [
    ("lp5_rb_noiseless", "lp", dict(include_noise=False)),
    ("lp5_rb_noisy", "lp", dict(include_noise=True)),
    ("lp5_cs_noisy", "lp", dict(include_noise=True, species="Cs133")),
    ("lp5_rb_noisy_bare_light_shifts", "lp", dict(include_noise=True, light_shift_compensation=False)),
    ("lp5_rb_noisy_far_detuned", "lp", dict(include_noise=True, Delta_e=2 * np.pi * 20e9, laser_1_power=50e-6 * 20,
                                            laser_2_power=0.3 * 20)),
    ("lp5_rb_noisy_weak_blockade", "lp", dict(include_noise=True, spacing_factor=5.373584,
                                              delta_over_omega=0.369547, omega_tau=4.215360)),
    ("bb5_rb_noisy", "jp", dict(include_noise=True)),
    ("bb5_cs_noiseless", "jp", dict(include_noise=False, species="Cs133")),
    ("bb7_rb_noisy", "jp", dict(include_noise=True, omega_tau=22.0,
                                switching_times=[2.0, 5.5, 9.0, 13.0, 16.5, 20.0],
                                phases=[0.3, -1.2, 2.0, 0.1, -2.4, 1.5, 0.7])),
    ("sjp5_rb_noisy_24steps", "sjp", dict(include_noise=True, n_smooth_steps=24)),
    ("lp5_rb_noisy_e_decays_to_0", "lp", dict(include_noise=True, intermediate_branching_1=0.0)),
]

ACT2 = [5 * a + b for a in (1, 2, 3) for b in (1, 2, 3)]     # both atoms in {1, e, r}
ACT01 = [1, 2, 3]                                           # atom A in |0>, atom B in {1, e, r}
ACT10 = [5, 10, 15]


def cplx(a):
    a = np.asarray(a)
    return [a.real.tolist(), a.imag.tolist()]


def main():
    out = []
    for name, proto, kw in CASES:
        plan = prologue.build_plan(proto, B=1, hilbert_space_dim=5, warn=False, **dict(MEDIUM, **kw))
        pt, seg = plan.points[0], plan.segs[0]
        noisy = bool(plan.include_noise)
        segs = [orc.Segment(s[0], complex(s[1], s[2]), s[3]) for s in seg if s[0] != 0.0]
        rates = orc.NoiseRates5(g_e1=pt[14], g_e0=pt[15], g_r1=pt[4], g_r0=pt[5], g_rl=pt[16], g_phi=pt[6]) if noisy else None
        fid, avg, info, fin = orc.evaluate5(segs, pt[0], pt[1], pt[12], pt[13], rates, ls_comp=pt[17])
        case = {"name": name, "protocol": proto, "noisy": noisy, "point": pt.tolist(), "segs": seg.tolist(),
                "expected": {"avg_fidelity": avg, "fidelities": fid,
                             "controlled_phase_rad": info["controlled_phase_rad"],
                             "cz_phase_fidelity": info["cz_phase_fidelity"]}}
        if noisy:
            r11 = fin["11"]
            case["expected"].update(
                rho11_active=cplx(r11[np.ix_(ACT2, ACT2)]), rho01_active=cplx(fin["01"][np.ix_(ACT01, ACT01)]),
                rho10_active=cplx(fin["10"][np.ix_(ACT10, ACT10)]),
                loss=float(1.0 - np.real(np.trace(r11[np.ix_(ACT2, ACT2)]))),
                trace=float(np.real(np.trace(r11))))
        else:
            case["expected"].update(psi11_active=cplx(fin["11"][ACT2]), psi01_active=cplx(fin["01"][ACT01]))
        out.append(case)
        print(f"{name:36s} F = {avg:.9f}   ctrl = {np.degrees(info['controlled_phase_rad']):8.3f} deg")
    path = os.path.join(HERE, "five_level_cases.json")
    with open(path, "w") as f:
        json.dump({"generator": "tests/golden/make_golden5.py", "oracle": "oracle/rydcz_oracle.py evaluate5 (expm)",
                   "cases": out}, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
