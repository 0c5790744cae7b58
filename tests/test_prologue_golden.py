"""The vectorised scalar prologue (rydberg_gates/prologue.py) against fixtures produced by the
reference's own prologue code (tests/golden/make_golden.py).  CPU only."""
import warnings

import numpy as np
import pytest

import _cases as C
from noisyquantumsimulator_b200.rydberg_gates import prologue

CASES = C.load("reference_cases.json") + C.load("known_answers.json")


def _rel(a, b):
    if a == b:          # covers +-inf (V/Ω = inf when Δe < 0 makes Ω negative)
        return 0.0
    return abs(a - b) / max(abs(b), 1e-300) if b != 0 else abs(a)


@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_scalars_rates_and_schedule(case):
    proto, kw = C.plan_kwargs(case)
    plan = prologue.build_plan(proto, B=1, warn=False, **kw)
    for k, v in case["scalars"].items():
        if k in ("xi_re", "xi_im", "n_pulses"):
            continue
        assert _rel(float(plan.scalars[k][0]), v) < 1e-12, k
    assert abs(plan.scalars["xi"][0] - complex(case["scalars"]["xi_re"], case["scalars"]["xi_im"])) < 1e-12
    assert plan.n_pulses == case["scalars"]["n_pulses"]
    for k, v in case["noise_breakdown"].items():
        got = plan.noise_breakdown[k]
        if isinstance(v, bool) or isinstance(v, str):
            assert got == v, k
        else:
            assert _rel(float(np.asarray(got, dtype=float).reshape(-1)[0]), v) < 1e-11, k
    s = case["segments"]
    want = np.stack([s["dt"], s["re"], s["im"], s["delta"]], 1)
    mine = plan.segs[0][plan.segs[0][:, 0] != 0]
    assert mine.shape == want.shape
    assert (np.abs(mine - want) / (np.abs(want).max(0) + 1e-300)).max() < 1e-12
    pt, _ = C.pack_case(case)
    assert np.allclose(plan.points[0, :9], pt[:9], rtol=1e-11, atol=0)


def test_batch_equals_one_by_one():
    """B points in one call == the same points one at a time (bitwise)."""
    rng = np.random.default_rng(3)
    B = 64
    kw = dict(laser_1_power=50e-6, laser_1_waist=50e-6, laser_2_power=10 ** rng.uniform(-1.3, 0.3, B),
              laser_2_waist=50e-6, Delta_e=2 * np.pi * 1e9, tweezer_power=0.02, tweezer_waist=0.8e-6,
              temperature=rng.uniform(1e-6, 5e-5, B), spacing_factor=rng.uniform(2.0, 6.0, B),
              delta_over_omega=rng.uniform(0.2, 0.5, B), omega_tau=rng.uniform(3.5, 5.5, B))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        plan = prologue.build_plan("lp", B=B, **kw)
        for i in (0, 17, 63):
            one = prologue.build_plan("lp", B=1, **{k: (v[i] if isinstance(v, np.ndarray) else v) for k, v in kw.items()})
            assert np.array_equal(one.points[0], plan.points[i])
            assert np.array_equal(one.segs[0], plan.segs[i])


def test_reference_warnings_and_errors():
    with pytest.warns(UserWarning, match="Blockade too weak"):
        prologue.build_plan("lp", B=1, laser_1_power=50e-6, laser_1_waist=50e-6, laser_2_power=0.3,
                            laser_2_waist=50e-6, Delta_e=2 * np.pi * 1e9, spacing_factor=6.5)
    with pytest.warns(UserWarning, match="very low"):
        prologue.build_plan("jp", B=1, laser_1_power=1e-9, laser_1_waist=50e-6, laser_2_power=1e-4,
                            laser_2_waist=50e-6, Delta_e=2 * np.pi * 1e9)
    with pytest.raises(ValueError, match="Unknown species"):
        prologue.build_plan("lp", B=1, laser_1_power=50e-6, laser_1_waist=50e-6, laser_2_power=0.3,
                            laser_2_waist=50e-6, Delta_e=1e9, species="K39")
    with pytest.raises(AssertionError, match="Need len"):
        prologue.build_plan("jp", B=1, laser_1_power=50e-6, laser_1_waist=50e-6, laser_2_power=0.3,
                            laser_2_waist=50e-6, Delta_e=2 * np.pi * 1e9, switching_times=[1.0, 2.0],
                            phases=[0.0, 1.0], omega_tau=5.0)
