"""The reference's optimiser / exploration callers, population-at-once.
CPU: the evaluator is the host harness; the GPU tests run the same entry points on the CUDA path.

Parity statement: a population objective value must equal the reference's per-candidate pipeline
(_build_*_inputs -> simulate_CZ_gate -> extract_metrics -> compute_cost, RG/optimize_cz_gate.py:1081-1150)
for the same parameters."""
import importlib
import json
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import _cases as C


@pytest.fixture()
def host_eval(monkeypatch):
    from noisyquantumsimulator_b200 import engine
    monkeypatch.setitem(engine._default, 0, C.HostsimEvaluator())


def test_reference_import_paths_resolve():
    from qpu_simulator.micro_physics.neutral_atoms.rydberg_gates.optimize_cz_gate import (  # noqa: F401
        ApparatusConstraints, SimulationCache, OptimizationResult, optimize_cz_gate, run_baseline, compute_cost,
        extract_metrics, warm_start_bounds, _build_lp_inputs, _build_jp_bangbang_inputs, _build_smooth_jp_inputs,
        _get_lp_bounds_and_x0, _get_jp_bangbang_bounds_and_x0, _get_smooth_jp_bounds_and_x0)
    from qpu_simulator.micro_physics.neutral_atoms.rydberg_gates.optimization import (  # noqa: F401
        HardwareOptimizationResult, optimize_CZ_parameters, EvaluatedPoint, ExplorationResult,
        explore_parameter_space, combine_explorations)
    from qpu_simulator.micro_physics.neutral_atoms import rydberg_gates as rg
    assert rg.optimize_cz_gate is optimize_cz_gate


def test_compute_cost_and_cache_semantics(tmp_path):
    ocg = importlib.import_module("noisyquantumsimulator_b200.rydberg_gates.optimize_cz_gate")
    m = dict(avg_fidelity=0.99, f11=0.98, cz_phase_fidelity=0.97)
    assert ocg.compute_cost(m, gate_time_us=0.4, time_weight=0.01) == pytest.approx(10 * 1 + 5 * 4 + 2 * 9 + 0.004)
    assert ocg.compute_cost(dict(m, avg_fidelity=0.49)) == 1e6
    assert ocg.compute_cost(dict(m, f11=float("nan"))) == 1e6
    assert ocg.compute_cost({}) == 1e6
    cache = ocg.SimulationCache(precision=4)
    k = cache.make_key("lp", [0.37737100001, 4.29268], "abc")
    assert k == "abc|lp|(0.3774, 4.2927)"
    assert k not in cache
    cache[k] = (1.5, {"avg_fidelity": 0.9})
    assert k in cache and cache[k] == (1.5, {"avg_fidelity": 0.9}) and cache.hits == 1 and len(cache) == 1
    cache.misses += 3
    assert cache.hit_rate == 0.25
    path = str(tmp_path / "cache.json")
    cache.save(path)
    assert set(json.load(open(path))) == {"precision", "entries"}
    other = ocg.SimulationCache(precision=2)
    other.load(path)
    assert other.precision == 4 and other[k][0] == 1.5
    other.load(str(tmp_path / "missing.json"))            # silently ignored, like the reference
    a = ocg.ApparatusConstraints()
    assert len(a.fingerprint()) == 12 and a.fingerprint() != ocg.ApparatusConstraints(spacing_factor=3.0).fingerprint()
    exc = a.make_excitation_config(pol_purity=0.99)
    assert exc.laser_2.polarization == "sigma+" and exc.laser_1.polarization_purity == 0.99
    assert not a.make_noiseless().include_doppler_dephasing and a.make_full_noise().intensity_noise_frac == 0.01


@pytest.mark.parametrize("proto,nseg", [("lp", 5), ("jp_bangbang", 5), ("jp_bangbang", 7), ("smooth_jp", 5)])
@pytest.mark.parametrize("noise", [False, True])
def test_population_objective_equals_per_candidate_pipeline(host_eval, proto, nseg, noise):
    from noisyquantumsimulator_b200 import rydberg_gates as rg
    ocg = importlib.import_module("noisyquantumsimulator_b200.rydberg_gates.optimize_cz_gate")
    app = ocg.ApparatusConstraints()
    p = ocg._norm(proto)
    bounds, x0, names = ocg.bounds_and_x0(p, nseg)
    rng = np.random.default_rng(5)
    lo, hi = np.array(bounds).T
    X = lo[:, None] + (hi - lo)[:, None] * rng.uniform(size=(len(bounds), 6))
    X[:, 0] = x0
    X[:, 5] = X[:, 4]                                     # a duplicate inside the generation
    nz = app.make_full_noise() if noise else app.make_noiseless()
    cache = ocg.SimulationCache()
    f = ocg.PopulationObjective(p, app, nz, noise, 0.01, nseg, cache)
    cost = f(X)
    assert f.generations == 1 and len(cache) == 5 and cost[4] == cost[5]
    exc = app.make_excitation_config(pol_purity=0.99 if noise else 1.0)
    for i in (0, 2, 4):
        if p == "lp":
            inp = ocg._build_lp_inputs(X[:, i], exc, nz)
        elif p == "jp":
            inp = ocg._build_jp_bangbang_inputs(X[:, i], exc, nz, n_segments=nseg)
        else:
            inp = ocg._build_smooth_jp_inputs(X[:, i], exc, nz)
        res = rg.simulate_CZ_gate(inp, species=app.species, n_rydberg=app.n_rydberg, spacing_factor=app.spacing_factor,
                                  temperature=app.temperature, tweezer_power=app.tweezer_power,
                                  tweezer_waist=app.tweezer_waist, B_field=app.B_field, NA=app.NA, include_noise=noise)
        m = ocg.extract_metrics(res)
        ref = ocg.compute_cost(m, gate_time_us=m["gate_time_us"], time_weight=0.01)
        assert cost[i] == pytest.approx(ref, rel=1e-8, abs=1e-9), (proto, i)
        got = cache[f._key(X[:, i])][1]
        for k in ("avg_fidelity", "f11", "cz_phase_fidelity", "gate_time_us", "V_over_Omega", "Omega_MHz"):
            assert got[k] == pytest.approx(m[k], rel=1e-9, abs=1e-10)
    hits = f.cache_hits
    again = f(X)                                          # everything is served from the cache now
    assert np.array_equal(again, cost) and f.generations == 1 and f.cache_hits == hits + 6
    assert f(X[:, 2]) == cost[2]                          # 1-D call (SciPy's polish) returns a float


def test_optimize_cz_gate_contract(host_eval):
    ocg = importlib.import_module("noisyquantumsimulator_b200.rydberg_gates.optimize_cz_gate")
    app = ocg.ApparatusConstraints(spacing_factor=5.373584)          # the notebook's optimised spacing, V/Ω = 6.9
    base = ocg.run_baseline("lp", app, verbose=False)
    r = ocg.optimize_cz_gate("lp", app, maxiter=8, popsize=8, verbose=False, cache=ocg.SimulationCache())
    assert isinstance(r, ocg.OptimizationResult) and r.protocol == "lp" and r.discrete_variant == "default"
    assert r.param_names == ["delta_over_omega", "omega_tau"] and r.best_params.shape == (2,)
    assert r.best_cost < ocg.compute_cost(base, base["gate_time_us"]) and r.best_metrics["avg_fidelity"] > 0.999
    assert r.n_generations < r.n_evaluations / 4 and "Optimal parameters" in repr(r)
    assert set(r.all_variants) == {"default"}
    b, x0 = ocg.warm_start_bounds(r, frac=0.2, original_bounds=[(0.2, 0.5), (3.5, 5.5)])
    assert all(lo < v < hi for (lo, hi), v in zip(b, x0)) and b[0][0] >= 0.2
    # spacing as an extra dimension; bang-bang variants
    r2 = ocg.optimize_cz_gate("lp", ocg.ApparatusConstraints(), optimize_spacing=True, spacing_bounds=(4.0, 6.0),
                              maxiter=3, popsize=6, verbose=False, cache=ocg.SimulationCache(), polish=False)
    assert r2.param_names[-1] == "spacing_factor" and 4.0 <= r2.best_params[-1] <= 6.0
    r3 = ocg.optimize_cz_gate("jp_bangbang", ocg.ApparatusConstraints(), maxiter=1, popsize=2, verbose=False,
                              cache=ocg.SimulationCache(), polish=False)
    assert set(r3.all_variants) == {"5-segment", "7-segment"} and r3.discrete_variant in r3.all_variants
    assert len(r3.best_params) in (10, 14)
    with pytest.raises(ValueError):
        ocg.optimize_cz_gate("jp", ocg.ApparatusConstraints(), variant="6-segment", verbose=False)
    with pytest.raises(ValueError):
        ocg.optimize_cz_gate("xy", ocg.ApparatusConstraints(), verbose=False)
    r4 = ocg.optimize_cz_gate("smooth_jp", ocg.ApparatusConstraints(), strategy="two_phase", maxiter=1, popsize=2,
                              verbose=False, cache=ocg.SimulationCache(), polish=False)
    assert r4.param_names[0] == "omega_tau" and np.isfinite(r4.best_cost)


def test_exploration_records_and_pareto_front():
    from noisyquantumsimulator_b200.rydberg_gates import optimization as opt
    res = opt.ExplorationResult(protocol="lp", species="Rb87")
    for F, t in ((0.90, 100.0), (0.95, 200.0), (0.93, 250.0), (0.99, 400.0), (0.985, 500.0)):
        res.add_point(opt.EvaluatedPoint(Omega_MHz=5, laser_linewidth_kHz=1, V_over_Omega=50, fidelity=F,
                                         gate_time_ns=t, infidelity=1 - F))
    res.compute_pareto_front()
    assert [p.gate_time_ns for p in res.pareto_front] == [100.0, 200.0, 400.0]
    assert res.get_best_for_target(target_fidelity=0.94).gate_time_ns == 200.0
    assert res.get_best_for_target(target_time_ns=300).fidelity == 0.95
    assert res.get_best_for_target(target_fidelity=0.999) is None and "Pareto front: 3 points" in res.summary()
    both = opt.combine_explorations(res, res)
    assert both.n_evaluations == 10 and len(both.pareto_front) == 3


def test_hardware_population_matches_single_calls(host_eval):
    """evaluate_hardware_population (grouped by n_rydberg) against one simulate_CZ_gate per candidate built
    the way RG/optimization.py:856-915 builds it."""
    from noisyquantumsimulator_b200 import rydberg_gates as rg
    from noisyquantumsimulator_b200.rydberg_gates import optimization as opt
    rng = np.random.default_rng(8)
    P = 7
    cols = dict(rydberg_power_2=rng.uniform(0.05, 0.5, P), rydberg_power_1=rng.uniform(20e-6, 80e-6, P),
                temperature=rng.uniform(1e-6, 15e-6, P), spacing_factor=rng.uniform(2.5, 5, P),
                n_rydberg=np.array([60.2, 70.4, 69.6, 80, 60, 71, 70]), tweezer_power=rng.uniform(10e-3, 100e-3, P),
                tweezer_waist=rng.uniform(0.5e-6, 2e-6, P), laser_linewidth=rng.uniform(100, 1e4, P),
                delta_over_omega=rng.uniform(0.32, 0.42, P), omega_tau=rng.uniform(3.9, 4.8, P))
    ev = opt.evaluate_hardware_population("lp", cols, species="Rb87", laser_1_waist=50e-6, laser_2_waist=50e-6)
    assert ev["ok"].all()
    for i in range(P):
        mk = lambda pw: rg.LaserParameters(power=pw, waist=50e-6, linewidth_hz=cols["laser_linewidth"][i])
        inp = rg.LPSimulationInputs(
            excitation=rg.TwoPhotonExcitationConfig(laser_1=mk(cols["rydberg_power_1"][i]), laser_2=mk(cols["rydberg_power_2"][i])),
            noise=rg.NoiseSourceConfig(include_motional_dephasing=True, include_doppler_dephasing=True,
                                       include_intensity_noise=True, intensity_noise_frac=0.01),
            delta_over_omega=cols["delta_over_omega"][i], omega_tau=cols["omega_tau"][i])
        r = rg.simulate_CZ_gate(inp, species="Rb87", n_rydberg=int(round(cols["n_rydberg"][i])),
                                temperature=cols["temperature"][i], spacing_factor=cols["spacing_factor"][i],
                                tweezer_power=cols["tweezer_power"][i], tweezer_waist=cols["tweezer_waist"][i],
                                include_noise=True, return_dataclass=False)
        assert ev["fidelity"][i] == pytest.approx(r["avg_fidelity"], abs=1e-10)
        assert ev["gate_time_ns"][i] == pytest.approx(r["tau_total"] * 1e9, rel=1e-12)
        assert ev["V_over_Omega"][i] == pytest.approx(r["V_over_Omega"], rel=1e-12)
        assert ev["gamma_phi_thermal"][i] == pytest.approx(r["noise_breakdown"]["gamma_phi_thermal"], rel=1e-10)


def test_explorer_and_inverse_optimiser_run(host_eval):
    from noisyquantumsimulator_b200.rydberg_gates import optimization as opt
    res = opt.explore_parameter_space("lp", maxiter=1, popsize=2, verbose=False,
                                      bounds_rydberg_power_2=(0.05, 0.5), bounds_rydberg_power_1=(20e-9, 80e-9))
    assert res.n_evaluations == len(res.points) > 0 and res.pareto_front and res.optimizer_settings["seeds"] == [42]
    assert all(0 <= p.fidelity <= 1 and p.gate_time_ns > 0 for p in res.points)
    h = opt.optimize_CZ_parameters(target_fidelity=0.97, target_gate_time_ns=400, maxiter=2, popsize=3, verbose=False,
                                   polish=False, fixed_params={"n_rydberg": 70})
    assert isinstance(h, opt.HardwareOptimizationResult) and h.optimal_parameters["n_rydberg"] == 70
    assert h.n_evaluations > 0 and 0 < h.achieved_fidelity <= 1 and "Target" in repr(h)
    with pytest.raises(ValueError):
        opt.optimize_CZ_parameters(protocol="nope", verbose=False)


# ---- N > 1: two gloo ranks share a generation -------------------------------------------------
def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import importlib
    import _cases as C2
    from noisyquantumsimulator_b200 import engine
    ocg = importlib.import_module("noisyquantumsimulator_b200.rydberg_gates.optimize_cz_gate")
    engine._default[0] = C2.HostsimEvaluator()
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r = ocg.optimize_cz_gate("lp", ocg.ApparatusConstraints(spacing_factor=5.373584), maxiter=3, popsize=5, verbose=False,
                             cache=ocg.SimulationCache(), population_shards=(rank, world), polish=False)
    q.put((rank, r.best_params, r.best_cost))
    dist.barrier()
    dist.destroy_process_group()


def test_population_sharded_over_two_gloo_ranks(host_eval):
    ocg = importlib.import_module("noisyquantumsimulator_b200.rydberg_gates.optimize_cz_gate")
    ref = ocg.optimize_cz_gate("lp", ocg.ApparatusConstraints(spacing_factor=5.373584), maxiter=3, popsize=5, verbose=False,
                               cache=ocg.SimulationCache(), polish=False)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {r: (x, c) for r, x, c in (q.get(timeout=180) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        assert np.array_equal(got[r][0], ref.best_params) and got[r][1] == ref.best_cost


def _failing_worker(rank, world, port, q):
    """Rank 1's evaluator raises inside the batch call: both ranks must still leave the all-gather, with the
    reference's failure cost 1e6 for that shard and identical costs on every rank."""
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import _cases as C2
    from noisyquantumsimulator_b200 import engine
    ev = C2.HostsimEvaluator()
    if rank == 1:
        def boom(*a, **k):
            raise RuntimeError("injected CUDA failure")
        ev.simulate_host_select = boom
        ev.simulate_host_cols = boom
    engine._default[0] = ev
    ocg = importlib.import_module("noisyquantumsimulator_b200.rydberg_gates.optimize_cz_gate")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a = ocg.ApparatusConstraints(spacing_factor=5.373584)
    f = ocg.PopulationObjective("lp", a, a.make_noiseless(), False, 0.01, 5, ocg.SimulationCache(),
                                population_shards=(rank, world))
    X = np.array([[0.36, 0.37, 0.38, 0.39], [4.1, 4.2, 4.3, 4.4]])
    q.put((rank, f(X)))
    dist.barrier()
    dist.destroy_process_group()


def test_failing_rank_does_not_hang_the_collective(host_eval):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_failing_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(got[0], got[1])
    assert (got[0][:2] < 1e5).all() and (got[0][2:] == 1e6).all()      # rank 0's shard evaluated, rank 1's failed


def test_selected_output_columns_equal_the_full_rows(host_eval):
    """outputs=... (rydcz_simulate_host_select on the GPU) returns the same numbers as the full result rows."""
    from noisyquantumsimulator_b200 import rydberg_gates as rg
    from noisyquantumsimulator_b200.rydberg_gates.simulation import METRIC_OUTPUTS
    ocg = importlib.import_module("noisyquantumsimulator_b200.rydberg_gates.optimize_cz_gate")
    a = ocg.ApparatusConstraints()
    kw = a.plan_kwargs(True, a.make_full_noise())
    rng = np.random.default_rng(5)
    ot = rng.uniform(5, 30, 6)
    sw = np.sort(rng.uniform(0.05, 0.95, (6, 4)), axis=1) * ot[:, None]
    ph = rng.uniform(-np.pi, np.pi, (6, 5))
    full = rg.simulate_CZ_gate_batch("jp", B=6, warn=False, omega_tau=ot, switching_times=sw, phases=ph, **kw)
    sel = rg.simulate_CZ_gate_batch("jp", B=6, warn=False, omega_tau=ot, switching_times=sw, phases=ph,
                                    outputs=METRIC_OUTPUTS + ("loss_probability",), **kw)
    assert sel.raw is None and set(sel.columns) == set(METRIC_OUTPUTS) | {"loss_probability"}
    for name in METRIC_OUTPUTS + ("loss_probability",):
        assert np.array_equal(sel.column(name), full.column(name)), name
        assert sel.column(name).flags.c_contiguous
    assert np.array_equal(sel.fidelities, full.fidelities) and np.array_equal(sel.cost(), full.cost())
    assert np.array_equal(ocg.batch_metric_rows(sel), ocg.batch_metric_rows(full))
    with pytest.raises(KeyError):
        sel.column("phi11")
    with pytest.raises(ValueError, match="unknown output"):
        rg.simulate_CZ_gate_batch("jp", B=6, warn=False, omega_tau=ot, outputs=("nonsense",), **kw)
