"""The device prologue (csrc/rydcz_prologue.cuh, compiled for the host by tests/hostsim) against
the NumPy prologue, which is itself pinned to the reference's prologue by test_prologue_golden.py."""
import warnings

import numpy as np
import pytest

import _cases as C
from noisyquantumsimulator_b200.rydberg_gates import device_prologue as dp
from noisyquantumsimulator_b200.rydberg_gates import prologue

# every fixture, shaped LP pulses included (998 envelope-sampled segments with the scalar-envelope quirk)
CASES = C.load("reference_cases.json") + C.load("known_answers.json")


def _close(a, b, tol=2e-12):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = np.maximum(np.abs(b), 1e-300)
    return np.all((a == b) | (np.abs(a - b) <= tol * scale))


@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_device_prologue_matches_numpy_prologue(case):
    proto, kw = C.plan_kwargs(case)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        plan = prologue.build_plan(proto, B=1, warn=False, **kw)
    mk, ak = dp.split_kwargs(kw)
    nseg = plan.segs.shape[1]
    model = dp.model_constants(proto, n_segments=nseg, **mk)
    app = dp.pack_apparatus(proto, 1, **ak)
    pts, segs, ns, sc = C.hostsim_prologue(model, app)
    assert ns[0] == nseg
    assert _close(pts[0], plan.points[0]), (pts[0], plan.points[0])
    scale = np.abs(plan.segs[0]).max(0) + 1e-300
    assert (np.abs(segs[0] - plan.segs[0]) / scale).max() < 5e-12
    s = plan.scalars
    want = [s["Omega"][0], s["V"][0], s["tau_total"][0], s["tau_single"][0], s["Delta"][0], s["xi"][0].real,
            s["xi"][0].imag, s["omega_tau"][0]]
    assert _close(sc[0], want, 1e-11)


def test_batch_rows_with_arrays():
    rng = np.random.default_rng(5)
    B = 257
    kw = dict(laser_1_power=50e-6, laser_1_waist=50e-6, laser_2_power=10 ** rng.uniform(-1.3, 0.3, B),
              laser_2_waist=50e-6, Delta_e=2 * np.pi * 1e9, tweezer_power=0.02, tweezer_waist=0.8e-6,
              temperature=rng.uniform(1e-6, 5e-5, B), spacing_factor=rng.uniform(2.0, 6.0, B),
              omega_tau=rng.uniform(5, 40, B), switching_times=np.sort(rng.uniform(1, 5, (B, 4)), 1),
              phases=rng.uniform(-3, 3, (B, 5)))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        plan = prologue.build_plan("jp", B=B, warn=False, **kw)
    mk, ak = dp.split_kwargs(kw)
    pts, segs, ns, sc = C.hostsim_prologue(dp.model_constants("jp", n_segments=5, **mk), dp.pack_apparatus("jp", B, **ak))
    assert _close(pts, plan.points)
    assert (np.abs(segs - plan.segs) / (np.abs(plan.segs).max((0, 1)) + 1e-300)).max() < 5e-12


@pytest.mark.parametrize("proto", ["lp", "jp", "sjp"])
@pytest.mark.parametrize("noise", [True, False])
def test_dim5_rows(proto, noise):
    """hilbert_space_dim=5: the 20-double rows (Omega_1, Delta_e, |e> decay split, |r> loss split, light-shift flag)."""
    rng = np.random.default_rng(9)
    B = 33
    kw = dict(laser_1_power=50e-6, laser_1_waist=50e-6, laser_2_power=10 ** rng.uniform(-1.3, 0.3, B),
              laser_2_waist=50e-6, Delta_e=2 * np.pi * rng.uniform(0.8e9, 5e9, B), tweezer_power=0.02,
              tweezer_waist=0.8e-6, temperature=rng.uniform(1e-6, 5e-5, B), spacing_factor=rng.uniform(2.0, 6.0, B),
              hilbert_space_dim=5, include_noise=noise, intermediate_branching_1=0.3, species="Cs133",
              light_shift_compensation=bool(noise))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        plan = prologue.build_plan(proto, B=B, warn=False, **kw)
    assert plan.points.shape == (B, 20)
    mk, ak = dp.split_kwargs(kw)
    model = dp.model_constants(proto, n_segments=plan.segs.shape[1], **mk)
    pts, segs, ns, sc = C.hostsim_prologue(model, dp.pack_apparatus(proto, B, **ak))
    assert pts.shape == (B, 20) and _close(pts, plan.points, 5e-12)
    assert (np.abs(segs - plan.segs) / (np.abs(plan.segs).max((0, 1)) + 1e-300)).max() < 5e-12


def test_shaped_pulses_through_the_batch_api(monkeypatch):
    """simulate_CZ_gate_batch routes gaussian / cosine / blackman LP pulses through the device prologue now; the
    results equal the host-prologue route, and the degenerate shapes (constant amplitude: the scalar-envelope quirk)
    qualify for the repeated-segment path (998 segments that differ only in the laser phase)."""
    from noisyquantumsimulator_b200 import engine
    from noisyquantumsimulator_b200 import rydberg_gates as rg
    fake = C.HostsimEvaluator()
    monkeypatch.setattr(engine, "get_evaluator", lambda device=0: fake)
    rng = np.random.default_rng(2)
    B = 3
    base = dict(laser_1_power=50e-6, laser_1_waist=50e-6, laser_2_power=10 ** rng.uniform(-0.8, 0.3, B), laser_2_waist=50e-6,
                Delta_e=2 * np.pi * 1e9, tweezer_power=0.02, tweezer_waist=0.8e-6, warn=False)
    for shape in ("gaussian", "cosine", "blackman"):
        a = rg.simulate_CZ_gate_batch("lp", B=B, pulse_shape=shape, prologue="device", **base)
        b = rg.simulate_CZ_gate_batch("lp", B=B, pulse_shape=shape, prologue="host", **base)
        assert np.abs(a.raw[:, :14] - b.raw[:, :14]).max() < 1e-9, shape
        c = rg.simulate_CZ_gate_batch("lp", B=B, pulse_shape=shape, outputs=("avg_fidelity", "tau_total"), **base)
        assert np.array_equal(c.avg_fidelity, a.avg_fidelity)
    with pytest.raises(TypeError):           # 'drag' lacks its Delta_leak argument in the reference as well
        rg.simulate_CZ_gate_batch("lp", B=B, pulse_shape="drag", **base)


def test_model_constants_are_memoised_per_keyword_set():
    """An optimiser calls with the same model thousands of times: the constants come from a cache, read-only, and a
    different keyword set (or an explicit AtomicConfiguration) is never answered from another set's entry."""
    from noisyquantumsimulator_b200.rydberg_gates.configurations import AtomicConfiguration
    a = dp.model_constants("jp", n_segments=5, species="Rb87", n_rydberg=70, include_noise=True)
    b = dp.model_constants("jp", n_segments=5, species="Rb87", n_rydberg=70, include_noise=True)
    assert a is b and not a.flags.writeable
    assert np.array_equal(a, dp._model_constants("jp", n_segments=5, species="Rb87", n_rydberg=70, include_noise=True))
    c = dp.model_constants("jp", n_segments=5, species="Cs133", n_rydberg=70, include_noise=True)
    d = dp.model_constants("jp", n_segments=7, species="Rb87", n_rydberg=70, include_noise=True)
    e = dp.model_constants("lp", species="Rb87", n_rydberg=70, include_noise=True)
    assert not np.array_equal(a, c) and d[2] == 7 and a[2] == 5 and e[1] == 0 and a[1] == 1
    cfg = AtomicConfiguration(species="Rb87", qubit_0=(1, 0), qubit_1=(2, 0), n_rydberg=60, L_rydberg="S")
    f = dp.model_constants("jp", n_segments=5, config=cfg)
    g = dp.model_constants("jp", n_segments=5, config=cfg)
    assert f is not g and np.array_equal(f, g) and f.flags.writeable and not np.array_equal(f, a)
