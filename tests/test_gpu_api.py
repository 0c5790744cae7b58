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


def test_prologue_kernel_matches_numpy_prologue():
    """prologue_kernel on the GPU against rydberg_gates/prologue.py (itself pinned to the reference's
    prologue by the golden fixtures): points, schedules and reported scalars for all three protocols."""
    import torch
    import warnings
    from noisyquantumsimulator_b200.engine import get_evaluator
    from noisyquantumsimulator_b200.rydberg_gates import device_prologue as dpro, prologue
    ev = get_evaluator(0)
    rng = np.random.default_rng(11)
    B = 4099
    base = dict(laser_1_power=50e-6, laser_1_waist=50e-6, laser_2_power=10 ** rng.uniform(-1.3, 0.3, B),
                laser_2_waist=50e-6, Delta_e=2 * np.pi * 1e9, tweezer_power=rng.uniform(0.01, 0.05, B),
                tweezer_waist=0.8e-6, temperature=rng.uniform(1e-6, 5e-5, B), spacing_factor=rng.uniform(2.0, 6.5, B),
                B_field=rng.uniform(1e-5, 5e-4, B))
    protos = {
        "lp": dict(delta_over_omega=rng.uniform(0.2, 0.5, B)),                  # omega_tau from the table
        "jp": dict(omega_tau=rng.uniform(5, 40, B), switching_times=np.sort(rng.uniform(1, 5, (B, 6)), 1),
                   phases=rng.uniform(-3, 3, (B, 7))),
        "sjp": dict(omega_tau=rng.uniform(5, 25, B), A=rng.uniform(0.2, 3, B), phi_offset=rng.uniform(0, 6, B)),
    }
    for proto, pk in protos.items():
        for dim in (3, 4):
            kw = dict(base, hilbert_space_dim=dim, **pk)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                plan = prologue.build_plan(proto, B=B, warn=False, **kw)
            mk, ak = dpro.split_kwargs(kw)
            S = plan.segs.shape[1]
            model = dpro.model_constants(proto, n_segments=S, **mk)
            d_app = torch.from_numpy(dpro.pack_apparatus(proto, B, **ak)).cuda()      # [40, B] slot-major
            d_pts = torch.empty((B, 12), dtype=torch.float64, device="cuda")
            d_seg = torch.empty((B, S, 4), dtype=torch.float64, device="cuda")
            d_ns = torch.empty(B, dtype=torch.int32, device="cuda")
            d_sc = torch.empty((B, 8), dtype=torch.float64, device="cuda")
            ev.prologue_device(model, d_app, d_pts, d_seg, d_ns, d_sc)
            torch.cuda.synchronize()
            pts, seg = d_pts.cpu().numpy(), d_seg.cpu().numpy()
            ok = (pts == plan.points) | (np.abs(pts - plan.points) <= 5e-12 * np.abs(plan.points))
            assert ok.all(), (proto, dim, np.argwhere(~ok)[:5])
            scale = np.abs(plan.segs).max((0, 1)) + 1e-300
            assert (np.abs(seg - plan.segs) / scale).max() < 1e-11
            assert np.array_equal(d_ns.cpu().numpy(), plan.n_seg)
            sc = d_sc.cpu().numpy()
            assert np.allclose(sc[:, 0], plan.scalars["Omega"], rtol=1e-13)
            assert np.allclose(sc[:, 2], plan.scalars["tau_total"], rtol=1e-13)


@pytest.mark.parametrize("proto", ["lp", "jp", "sjp"])
def test_batch_device_prologue_equals_host_prologue(proto):
    from noisyquantumsimulator_b200 import rydberg_gates as rg
    rng = np.random.default_rng(21)
    B = 300
    kw = dict(laser_1_power=50e-6, laser_1_waist=50e-6, laser_2_power=10 ** rng.uniform(-1.0, 0.3, B),
              laser_2_waist=50e-6, Delta_e=2 * np.pi * 1e9, tweezer_power=0.02, tweezer_waist=0.8e-6,
              temperature=rng.uniform(1e-6, 3e-5, B), spacing_factor=rng.uniform(2.5, 5.0, B), warn=False)
    if proto == "jp":
        kw.update(omega_tau=rng.uniform(5, 40, B), switching_times=np.sort(rng.uniform(1, 5, (B, 4)), 1),
                  phases=rng.uniform(-3, 3, (B, 5)))
    a = rg.simulate_CZ_gate_batch(proto, B=B, prologue="device", return_states=True, **kw)
    b = rg.simulate_CZ_gate_batch(proto, B=B, prologue="host", return_states=True, **kw)
    assert np.abs(a.raw[:, :14] - b.raw[:, :14]).max() < 1e-9
    assert np.abs(a.states - b.states).max() < 1e-9
    assert np.allclose(a.tau_total, b.tau_total, rtol=1e-13) and np.allclose(a.V_over_Omega, b.V_over_Omega, rtol=1e-13)
    assert np.allclose(a.cost(), b.cost(), rtol=1e-7, atol=1e-9)


def test_population_objective_and_sharded_evaluation_on_gpu():
    """Config 4 in miniature: a DE generation as one launch; the sharded evaluator with world = 1."""
    from noisyquantumsimulator_b200 import sharding
    from noisyquantumsimulator_b200.engine import get_evaluator
    from noisyquantumsimulator_b200.rydberg_gates import batched_objective as bo, prologue
    rng = np.random.default_rng(42)                      # the reference's DE seed (RG/optimize_cz_gate.py:797)
    for proto, nseg in (("lp", 5), ("jp", 5), ("sjp", 5)):
        bounds, x0, _ = bo.bounds_and_x0(proto, nseg)
        lo, hi = np.array(bounds).T
        X = lo[:, None] + (hi - lo)[:, None] * rng.uniform(size=(len(bounds), 1024))
        f = bo.make_vectorized_objective(proto, include_noise=True, n_segments=nseg)
        cost = f(X)
        assert cost.shape == (1024,) and np.all(np.isfinite(cost)) and cost.min() > 0
        one = f(X[:, 17])
        assert abs(one - cost[17]) <= 1e-9 * max(1.0, abs(one))
    res = bo.optimize_cz_gate_batched("lp", bo.ApparatusConstraints(spacing_factor=5.373584), maxiter=10, popsize=12)
    assert res.fun < 0.05
    kw = bo.ApparatusConstraints().plan_kwargs(True)
    plan = prologue.build_plan("jp", B=64, warn=False, omega_tau=rng.uniform(5, 40, 64), **kw)
    ev = get_evaluator(0)
    full = sharding.evaluate_sharded(ev, 3, True, plan.points, plan.segs, plan.n_seg, rank=0, world=1)
    assert np.array_equal(full, ev.evaluate_host(3, True, plan.points, plan.segs, plan.n_seg)[0])


def test_pipelined_column_entry_equals_block_entry():
    """rydcz_simulate_host_cols (varying slots only over PCIe, chunks pipelined on eight streams)
    against rydcz_simulate_host (whole [40, B] block, one stream): bit-identical rows, whatever the
    chunk size, with strided columns, with and without Monte-Carlo shots."""
    from noisyquantumsimulator_b200.engine import Evaluator
    from noisyquantumsimulator_b200.rydberg_gates import device_prologue as dpro
    ev = Evaluator(0)
    rng = np.random.default_rng(5)
    B = 9001
    kw = dict(laser_1_power=50e-6, laser_1_waist=50e-6, laser_2_power=10 ** rng.uniform(-1.0, 0.3, B),
              laser_2_waist=50e-6, Delta_e=2 * np.pi * 1e9, tweezer_power=0.02, tweezer_waist=0.8e-6,
              temperature=rng.uniform(1e-6, 3e-5, B), spacing_factor=2.8,
              omega_tau=rng.uniform(5, 40, B), switching_times=np.sort(rng.uniform(1, 5, (B, 4)), 1),
              phases=rng.uniform(-3, 3, (B, 5)))
    mk, ak = dpro.split_kwargs(kw)
    model = dpro.model_constants("jp", n_segments=5, **mk)
    cols = dpro.apparatus_columns("jp", B, **ak)
    assert sum(isinstance(c, np.ndarray) for c in cols) == 12 and cols[dpro.APP_PROTO + 2].strides == (32,)
    block = dpro.pack_apparatus("jp", B, **ak)
    ref = [np.array(a) for a in ev.simulate_host(model, block, want_states=True)]
    for chunk in (1 << 20, 2048, 1024):
        ev.set_chunk(chunk)
        got = ev.simulate_host_cols(model, cols, B, want_states=True)
        for a, b in zip(ref, got):
            assert np.array_equal(a, b), chunk
    # shots: the Philox counter carries the point's index in the whole batch
    ev.set_chunk(1 << 20)
    one = np.array(ev.simulate_host_cols(model, cols, B, n_shots=4, seed=77)[0])
    ev.set_chunk(4096)                                   # 1024 points x 4 shots per chunk
    many = np.array(ev.simulate_host_cols(model, cols, B, n_shots=4, seed=77)[0])
    assert np.array_equal(one, many)
    blk = np.array(ev.simulate_host(model, block, n_shots=4, seed=77)[0])
    assert np.array_equal(one, blk)
    ev.close()


def test_optimize_cz_gate_reaches_the_notebook_optimum():
    """The reference's LP optimisation at spacing 5.373584 (cz_gate_optimization_demo.ipynb cell 10:
    F = 0.999918 at delta/Omega = 0.369547, Omega*tau = 4.215360) -- population-at-once on the GPU."""
    import importlib
    ocg = importlib.import_module("noisyquantumsimulator_b200.rydberg_gates.optimize_cz_gate")
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = ocg.optimize_cz_gate("lp", ocg.ApparatusConstraints(spacing_factor=5.373584), maxiter=40, popsize=15,
                                 verbose=False, cache=ocg.SimulationCache())
    assert r.best_metrics["avg_fidelity"] >= 0.99990 and r.success
    assert abs(r.best_params[0] - 0.369547) < 0.01 and abs(r.best_params[1] - 4.215360) < 0.05
    assert r.n_generations <= 60 and r.n_evaluations > 1000


def test_batched_sweep_example_runs(tmp_path):
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "sweeps.json")
    subprocess.check_call([sys.executable, os.path.join(root, "examples", "research_parameter_sweeps_batched.py"),
                           "--points", "32", "--out", out], stdout=subprocess.DEVNULL)
    res = json.load(open(out))
    assert len(res["temperature_uK"]["lp"]["fidelity"]) == 32 and set(res["species"]) == {"Rb87", "Cs133"}
