"""GPU parity beyond the golden fixtures (round 2): 4-level and unequal-atom random batches against the expm
oracle, mixed-kind batches (the repeated-segment + clustered-general ordering), loss / Rydberg outputs, the
status column, column-selected returns, stream ordering, the registered torch operators, the evolve_* helpers,
and a reduced path-against-path stress.  Everything goes through the C ABI (include/rydcz.h)."""
import numpy as np
import pytest

import _cases as C
from oracle import rydcz_oracle as orc

pytestmark = pytest.mark.gpu

REF = {c["name"]: c for c in C.load("reference_cases.json")}


def _rates(rng, dim):
    return orc.NoiseRates(gamma_r=rng.uniform(5e3, 9e3), gamma_phi_laser=rng.uniform(1e2, 5e4),
                          gamma_phi_thermal=rng.uniform(0, 5e4), gamma_loss_antitrap=rng.uniform(0, 3e3),
                          gamma_loss_background=1e3, gamma_scatter_intermediate=rng.uniform(1e3, 2e4),
                          gamma_leakage=rng.uniform(0, 50), mJ_leakage_rate=(rng.uniform(10, 500) if dim == 4 else 0.0))


def _schedule(rng, protocol, Om, n_sjp=40):
    if protocol == "lp":
        return orc.segments_lp(Om, rng.uniform(0.2, 0.5) * Om, rng.uniform(3.5, 5.5) / Om)
    if protocol == "bb":
        ot = rng.uniform(5, 40)
        return orc.segments_bangbang(Om, ot, list(np.sort(rng.uniform(0.01, 0.99, 4)) * ot),
                                     list(rng.uniform(-np.pi, np.pi, 5)))
    return orc.segments_smooth_jp(Om, rng.uniform(5, 25) / Om, rng.uniform(0.05, 1) * np.pi, rng.uniform(0.5, 3) * Om,
                                  rng.uniform(0, 2 * np.pi), -rng.uniform(0.001, 0.1) * Om, n_sjp)


def _pack(dim, items):
    """items: [(V, d1, rates, (off_a, off_b), segments)]"""
    S = max(len(it[4]) for it in items)
    P = np.stack([C.pack_point(V, d1, r.collapsed(), off) for V, d1, r, off, _ in items])
    G = np.zeros((len(items), S, 4))
    for i, it in enumerate(items):
        G[i, :len(it[4])] = [[s.dt, s.Omega.real, s.Omega.imag, s.Delta] for s in it[4]]
    return P, G, np.array([len(it[4]) for it in items], dtype=np.int32)


def _compare(dim, noisy, items, out, st, tol_rho=1e-10, tol_f=1e-9):
    for i, (V, d1, r, off, sg) in enumerate(items):
        fid, avg, info, fin = orc.evaluate(dim, sg, V, d1, r if noisy else None, "exact", "coherence", off)
        ops = C.states_to_operators(dim, st[i])
        for lab in ("01", "10", "11"):
            got = ops["rho" + lab] if noisy else ops["coh" + lab][:, 0]
            assert np.abs(got - fin[lab]).max() < tol_rho, (i, lab)
        assert abs(out[i, C.OUT["avg"]] - avg) < tol_f
        for lab, slot in (("01", "f01"), ("10", "f10"), ("11", "f11")):
            assert abs(out[i, C.OUT[slot]] - fid[lab]) < tol_f
        assert abs(out[i, C.OUT["czf"]] - info["cz_phase_fidelity"]) < tol_f
        if noisy:
            # loss = P(an atom ended in |0>), ryd = P(an atom left in a Rydberg level), |11> input
            rho = fin["11"]
            n = dim * dim
            lvl = lambda k: (k // dim, k % dim)
            loss = sum(rho[k, k].real for k in range(n) if 0 in lvl(k))
            ryd = sum(rho[k, k].real for k in range(n) if any(x >= 2 for x in lvl(k)))
            assert abs(out[i, C.OUT["loss"]] - loss) < 1e-10 and abs(out[i, C.OUT["ryd"]] - ryd) < 1e-10


@pytest.mark.parametrize("protocol", ["lp", "bb"])
@pytest.mark.parametrize("noisy", [True, False])
def test_random_four_level_batch_against_oracle(evaluator, protocol, noisy):
    """The dim-4 model (r+ / r- with mJ mixing) on random points: folded plain and clustered paths."""
    rng = np.random.default_rng(4321)
    items = []
    for _ in range(6 if noisy else 12):      # the dense 256x256 expm oracle is slow
        Om = 2 * np.pi * rng.uniform(1.5e6, 9e6)
        items.append((Om * 10 ** rng.uniform(0.8, 2.4), 2 * np.pi * rng.uniform(0, 3e5), _rates(rng, 4), (0.0, 0.0),
                      _schedule(rng, protocol, Om)))
    P, G, ns = _pack(4, items)
    out, st = evaluator.evaluate_host(4, noisy, P, G, ns, want_states=True)
    _compare(4, noisy, items, out, st)


@pytest.mark.parametrize("dim", [3, 4])
@pytest.mark.parametrize("noisy", [True, False])
def test_random_unequal_atoms_against_oracle(evaluator, dim, noisy):
    """Atoms that see different detunings (what a Doppler shot is): general systems, plain and clustered."""
    rng = np.random.default_rng(99 + dim)
    items = []
    n = (16 if dim == 3 else 5) if noisy else (24 if dim == 3 else 10)
    for k in range(n):
        Om = 2 * np.pi * rng.uniform(1.5e6, 9e6)
        off = (2 * np.pi * rng.normal(0, 80e3), 2 * np.pi * rng.normal(0, 80e3))
        items.append((Om * 10 ** rng.uniform(0.8, 2.4), 2 * np.pi * rng.uniform(0, 3e5), _rates(rng, dim), off,
                      _schedule(rng, "lp" if k % 2 else "bb", Om)))
    P, G, ns = _pack(dim, items)
    out, st = evaluator.evaluate_host(dim, noisy, P, G, ns, want_states=True)
    _compare(dim, noisy, items, out, st)
    if dim == 3:
        status = evaluator.point_status(len(items)) & 0x0f
        # general plain / general clustered; without noise every system is a closed form and nothing is planned
        assert set(status) <= {0, 2} and ((status == 2).any() if noisy else (status == 0).all())


def test_mixed_kinds_in_one_sorted_batch(evaluator):
    """>= 64 points mixing unequal-atom strong-blockade points (clustered general kernel) with 300-step smooth-JP
    points (repeated-segment kernels), folded clustered and plain points: every row equals the single-interval
    path, and the repeated-segment points are really evaluated (they must sort in front of everything else)."""
    rng = np.random.default_rng(5)
    items = []
    for k in range(96):
        Om = 2 * np.pi * rng.uniform(2e6, 8e6)
        kind = k % 4
        if kind == 0:      # unequal atoms, strong blockade
            it = (Om * rng.uniform(150, 400), 2 * np.pi * 2e5, _rates(rng, 3),
                  (2 * np.pi * rng.normal(0, 50e3), 2 * np.pi * rng.normal(0, 50e3)), _schedule(rng, "bb", Om))
        elif kind == 1:    # smooth JP, 300 equal steps
            it = (Om * rng.uniform(20, 400), 2 * np.pi * 2e5, _rates(rng, 3), (0.0, 0.0), _schedule(rng, "sjp", Om, 300))
        elif kind == 2:    # identical atoms, strong blockade
            it = (Om * rng.uniform(150, 400), 2 * np.pi * 2e5, _rates(rng, 3), (0.0, 0.0), _schedule(rng, "lp", Om))
        else:              # weak blockade: plain folded
            it = (Om * rng.uniform(4, 9), 2 * np.pi * 2e5, _rates(rng, 3), (0.0, 0.0), _schedule(rng, "bb", Om))
        items.append(it)
    P, G, ns = _pack(3, items)
    for noisy in (True, False):
        evaluator.repeat_stats()
        out, st = evaluator.evaluate_host(3, noisy, P, G, ns, want_states=True)
        status = evaluator.point_status(len(items)) & 0x0f
        if noisy:
            assert evaluator.repeat_stats() == 24
            assert (status[1::4] == 5).all() and (status[0::4] == 2).all() and (status[2::4] == 3).all()
        else:      # without noise every system is a closed form: plain kernels (status 1 folded, 0 general), nothing planned
            assert evaluator.repeat_stats() == 0
            assert (status[1::4] == 1).all() and (status[0::4] == 0).all() and (status[2::4] == 1).all()
        assert np.isfinite(out).all()
        evaluator.set_cluster(0)
        out0, st0 = evaluator.evaluate_host(3, noisy, P, G, ns, want_states=True)
        evaluator.set_cluster(1)
        assert np.abs(out[:, :14] - out0[:, :14]).max() < 2e-11
        assert np.abs(st - st0).max() < 1e-11
    _compare(3, True, items[:8:2] + items[2:6:4], *[a[[0, 2, 4, 6, 2]] for a in
                                                    evaluator.evaluate_host(3, True, P, G, ns, want_states=True)])


def test_status_flags_and_five_level_general_rows(evaluator):
    """dim 5, a row with unequal atoms among identical ones: round 1 returned a NaN row and flagged the call; now the
    general systems evaluate it (status: plain), on the host and on the device entry alike, and no flag is raised."""
    import torch
    cases = [c for c in C.load("five_level_cases.json")][:3]
    P, G, ns = C.pack_batch5(cases)
    out, _ = evaluator.evaluate_host(5, True, P, G, ns)
    assert np.isfinite(out).all() and evaluator.call_flags() == (0, 0)
    odd = P.copy()
    odd[1, 3] = odd[1, 2] + 1.0e3                      # OFF_B != OFF_A
    oh, _ = evaluator.evaluate_host(5, True, odd, G, ns)
    dev = torch.device("cuda:0")
    o, _ = evaluator.evaluate_device(5, True, torch.from_numpy(odd).to(dev), torch.from_numpy(G).to(dev),
                                     torch.from_numpy(ns).to(dev))
    o = o.cpu().numpy()
    assert np.isfinite(o).all() and np.array_equal(o[:, :14], oh[:, :14])
    assert np.array_equal(o[[0, 2], :14], out[[0, 2], :14])
    assert 1e-12 < np.abs(o[1, :14] - out[1, :14]).max() < 1e-3       # 1 kHz of differential detuning: a small change
    assert evaluator.call_flags() == (0, 0)
    assert list(evaluator.point_status(3)) == [1, 0, 1]


def test_selected_columns_equal_full_rows_on_device(evaluator):
    from noisyquantumsimulator_b200 import rydberg_gates as rg
    from noisyquantumsimulator_b200.rydberg_gates.simulation import METRIC_OUTPUTS
    ocg = rg.optimize_cz_gate.__globals__
    a = rg.ApparatusConstraints()
    kw = a.plan_kwargs(True, a.make_full_noise())
    rng = np.random.default_rng(5)
    B = 70000                                            # three pipeline chunks, the last one ragged
    ot = rng.uniform(5, 30, B)
    sw = np.sort(rng.uniform(0.05, 0.95, (B, 4)), axis=1) * ot[:, None]
    ph = rng.uniform(-np.pi, np.pi, (B, 5))
    full = rg.simulate_CZ_gate_batch("jp", B=B, warn=False, omega_tau=ot, switching_times=sw, phases=ph, **kw)
    names = METRIC_OUTPUTS + ("loss_probability", "omega_tau")
    sel = rg.simulate_CZ_gate_batch("jp", B=B, warn=False, omega_tau=ot, switching_times=sw, phases=ph, outputs=names, **kw)
    for name in METRIC_OUTPUTS + ("loss_probability",):
        assert np.array_equal(sel.column(name), full.column(name)), name
    assert np.allclose(sel.column("omega_tau"), ot, rtol=1e-15)
    assert np.array_equal(ocg["batch_metric_rows"](sel), ocg["batch_metric_rows"](full))


def test_calls_on_different_streams_are_ordered(evaluator):
    """Two asynchronous calls of ONE context on two streams share the planning scratch: the second must wait for
    the first on the device (include/rydcz.h, 'Streams')."""
    import torch
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(11)
    batches = []
    for n in (40000, 3000):
        items = []
        Om = 2 * np.pi * 3.6e6
        pt = C.pack_point(Om * 342.5, 2 * np.pi * 2e5, _rates(rng, 3).collapsed())
        P = np.repeat(pt[None], n, 0)
        ot = rng.uniform(5, 40, n)
        G = np.zeros((n, 5, 4))
        fr = np.concatenate([np.zeros((n, 1)), np.sort(rng.uniform(0.01, 0.99, (n, 4)), axis=1), np.ones((n, 1))], axis=1)
        ph = rng.uniform(-np.pi, np.pi, (n, 5))
        G[:, :, 0] = np.diff(fr, axis=1) * ot[:, None] / Om
        G[:, :, 1], G[:, :, 2] = Om * np.cos(ph), Om * np.sin(ph)
        batches.append((torch.from_numpy(P).to(dev), torch.from_numpy(G).to(dev)))
    ref = [evaluator.evaluate_device(3, True, p, g)[0].clone() for p, g in batches]
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        o1, _ = evaluator.evaluate_device(3, True, *batches[0], stream=s1.cuda_stream)
        o2, _ = evaluator.evaluate_device(3, True, *batches[1], stream=s2.cuda_stream)
        torch.cuda.synchronize()
        assert torch.equal(o1, ref[0]) and torch.equal(o2, ref[1])


def test_torch_ops_match_the_c_abi(evaluator):
    import torch
    from noisyquantumsimulator_b200 import torch_ops  # noqa: F401
    from noisyquantumsimulator_b200.rydberg_gates import device_prologue as dpro
    dev = torch.device("cuda:0")
    cases = [c for c in REF.values() if c["dim"] == 3 and c["rates"] is not None][:12]
    pts, seg, ns = C.pack_batch(cases)
    out_h, st_h = evaluator.evaluate_host(3, True, pts, seg, ns, want_states=True)
    out, st = torch.ops.rydcz.evaluate(torch.from_numpy(pts).to(dev), torch.from_numpy(seg).to(dev),
                                       torch.from_numpy(ns).to(dev), 3, True, True)
    assert np.array_equal(out.cpu().numpy(), out_h) and np.array_equal(st.cpu().numpy(), st_h)
    status = torch.ops.rydcz.point_status(len(cases))
    assert status.dtype == torch.int32 and status.shape == (len(cases),)
    mean, shots = torch.ops.rydcz.evaluate_shots(torch.from_numpy(pts).to(dev), torch.from_numpy(seg).to(dev),
                                                 torch.from_numpy(ns).to(dev), 3, True, 4, 1234, True)
    mean_h, shots_h = evaluator.evaluate_shots_host(3, True, 4, 1234, pts, seg, ns, want_shots=True)
    assert np.array_equal(mean.cpu().numpy(), mean_h) and np.array_equal(shots.cpu().numpy(), shots_h)
    # simulate: device prologue + evaluate
    import bench
    cand = bench.make_candidates(2048, 77)
    mk, ak = dpro.split_kwargs(dict(bench.MEDIUM, **cand))
    model = dpro.model_constants("jp", n_segments=5, **mk)
    app = dpro.pack_apparatus("jp", 2048, **ak)
    o2, sc2 = torch.ops.rydcz.simulate(torch.from_numpy(model), torch.from_numpy(app).to(dev), 0, 1234)
    o_h, sc_h, _ = evaluator.simulate_host(model, app)
    assert np.array_equal(o2.cpu().numpy(), np.array(o_h)) and np.array_equal(sc2.cpu().numpy(), np.array(sc_h))
    with pytest.raises(RuntimeError, match="points must be"):
        torch.ops.rydcz.evaluate(torch.zeros(2, 11, dtype=torch.float64, device=dev),
                                 torch.zeros(2, 1, 4, dtype=torch.float64, device=dev), None, 3, True, False)


def test_evolve_helpers_on_the_cuda_path():
    from noisyquantumsimulator_b200 import rydberg_gates as rg
    from noisyquantumsimulator_b200.rydberg_gates import simulation as sim
    case = REF["bb5_noisy"]
    hs = rg.build_hilbert_space(3)
    one = {"0": hs.basis["0"], "1": hs.basis["1"]}
    kets = {a + b: np.kron(one[a], one[b]) for a in "01" for b in "01"}
    nb, sc, api = case["noise_breakdown"], case["scalars"], case["api"]["params"]
    c_ops, _ = rg.build_all_noise_operators(
        hs=hs, **{k: nb[k] for k in ("gamma_r", "gamma_bbr", "gamma_phi_laser", "gamma_phi_thermal", "gamma_phi_zeeman",
                                     "gamma_loss_antitrap", "gamma_loss_background", "gamma_scatter_intermediate",
                                     "gamma_leakage", "mJ_leakage_rate")})
    final, _ = sim.evolve_bangbang_jp(kets, sc["Omega"], sc["V"], api["omega_tau"], api["switching_times"], api["phases"],
                                      hs, c_ops, delta_zeeman=sc["delta_zeeman"], delta_stark=sc["delta_stark"])
    fin = C.golden_finals(case)
    for lab in ("00", "01", "10", "11"):
        assert np.abs(np.asarray(final[lab]) - fin[lab]).max() < 1e-10


def test_paths_agree_on_random_batches(evaluator):
    """A reduced tools/stress_paths.py inside the suite: planned paths (clustered, clustered general, repeated-segment)
    against the single-interval path on seeded batches over the SURVEY 8d ranges, 3- and 4-level, both branches."""
    rng = np.random.default_rng(2024)
    worst = 0.0
    for dim, noisy, n in ((3, True, 6000), (3, False, 6000), (4, True, 2000), (4, False, 2000)):
        Om = 2 * np.pi * rng.uniform(1.5e6, 9e6, n)
        V = Om * 10 ** rng.uniform(0.6, 2.9, n)
        P = np.zeros((n, 12))
        P[:, 0], P[:, 1] = V, 2 * np.pi * rng.uniform(0, 3e5, n)
        uneq = (rng.random(n) < 0.3) & (dim == 3)
        P[:, 2] = np.where(uneq, 2 * np.pi * rng.normal(0, 60e3, n), 0.0)
        P[:, 3] = np.where(uneq, 2 * np.pi * rng.normal(0, 60e3, n), P[:, 2])
        P[:, 4], P[:, 5] = rng.uniform(2e3, 5e3, n), rng.uniform(3e3, 3e5, n)
        P[:, 6], P[:, 7] = rng.uniform(1e2, 1e5, n), rng.uniform(1e3, 3e5, n)
        if dim == 4:
            P[:, 8] = rng.uniform(10, 500, n)
        S = 7
        G = np.zeros((n, S, 4))
        nseg = rng.choice([2, 5, 7], n)
        ot = rng.uniform(4, 40, n)
        fr = np.sort(rng.uniform(0.01, 0.99, (n, S - 1)), axis=1)
        ph = rng.uniform(-np.pi, np.pi, (n, S))
        for i in range(n):
            k = nseg[i]
            if k == 2:
                G[i, :2, 0] = 0.5 * ot[i] / Om[i]
                G[i, :2, 3] = rng.uniform(0.2, 0.5) * Om[i]
            else:
                b = np.concatenate([[0.0], np.sort(fr[i, :k - 1]), [1.0]])
                G[i, :k, 0] = np.diff(b) * ot[i] / Om[i]
            G[i, :k, 1], G[i, :k, 2] = Om[i] * np.cos(ph[i, :k]), Om[i] * np.sin(ph[i, :k])
        ns = nseg.astype(np.int32)
        out, st = evaluator.evaluate_host(dim, noisy, P, G, ns, want_states=True)
        planned, back = evaluator.cluster_stats()
        evaluator.set_cluster(0)
        out0, st0 = evaluator.evaluate_host(dim, noisy, P, G, ns, want_states=True)
        evaluator.set_cluster(1)
        # (3-level, noiseless: every system is a closed form on the plain kernels; nothing is planned)
        assert (planned == 0 if (dim == 3 and not noisy) else planned > n // 4) and back == 0
        worst = max(worst, np.abs(out[:, :14] - out0[:, :14]).max(), np.abs(st - st0).max())
    assert worst < 5e-11


def test_shaped_lp_pulses_on_the_device_prologue(evaluator):
    """Shaped LP pulses (998 envelope-sampled segments, scalar-envelope quirk) generated by prologue_kernel: equal to
    the host-prologue route and to the reference-run fixtures; the shapes that degenerate to a constant amplitude
    (gaussian, blackman) take the repeated-segment path, the cosine envelope the plain one."""
    from noisyquantumsimulator_b200 import rydberg_gates as rg
    from noisyquantumsimulator_b200.engine import get_evaluator
    rng = np.random.default_rng(2)
    B = 96
    base = dict(laser_1_power=50e-6, laser_1_waist=50e-6, laser_2_power=10 ** rng.uniform(-0.8, 0.3, B), laser_2_waist=50e-6,
                Delta_e=2 * np.pi * 1e9, tweezer_power=0.02, tweezer_waist=0.8e-6, warn=False)
    for shape, want in (("gaussian", 5), ("cosine", 1), ("blackman", 5)):
        for noise in (True, False):
            a = rg.simulate_CZ_gate_batch("lp", B=B, pulse_shape=shape, prologue="device", include_noise=noise, **base)
            status = get_evaluator(0).point_status(B) & 0x0f     # (the context simulate_CZ_gate_batch runs on)
            b = rg.simulate_CZ_gate_batch("lp", B=B, pulse_shape=shape, prologue="host", include_noise=noise, **base)
            assert np.abs(a.raw[:, :14] - b.raw[:, :14]).max() < 1e-9, (shape, noise)
            # (without noise every system of identical atoms has a closed form: plain folded path, no planning)
            assert (status == (want if noise else 1)).all(), (shape, noise, np.unique(status))
    # the fixtures generated by the reference's own evolve_shaped_pulse
    for name in ("lp_shaped_gaussian", "lp_shaped_cosine", "lp_shaped_blackman", "lp_shaped_cosine_noisy"):
        case = REF[name]
        proto, kw = C.plan_kwargs(case)
        r = rg.simulate_CZ_gate_batch(proto, B=1, prologue="device", return_states=True, warn=False, **kw)
        gold = case["result_coherence"] if case["rates"] is not None else case["result"]
        assert abs(r.avg_fidelity[0] - gold["avg_fidelity"]) < 1e-9, name
        fin = C.golden_finals(case)
        ops = C.states_to_operators(case["dim"], r.states[0])
        for lab in ("01", "10", "11"):
            got = ops["rho" + lab] if case["rates"] is not None else ops["coh" + lab][:, 0]
            assert np.abs(got - fin[lab]).max() < 1e-10, (name, lab)


def test_batch_compat_epilogue_on_the_gpu():
    """simulate_CZ_gate_batch(phase_mode="compat") against simulate_CZ_gate(phase_mode="compat") point by point
    (the reference's dominant-eigenvector formula, RG/simulation.py:424-434, on the rho's the kernel returns)."""
    from noisyquantumsimulator_b200 import rydberg_gates as rg
    from noisyquantumsimulator_b200.rydberg_gates.configurations import (LPSimulationInputs, LaserParameters,
                                                                         TwoPhotonExcitationConfig)
    from noisyquantumsimulator_b200.rydberg_gates.simulation import _plan_kwargs
    p2 = np.array([0.2, 0.35, 0.6, 1.1])

    def inputs(p):
        return LPSimulationInputs(excitation=TwoPhotonExcitationConfig(
            laser_1=LaserParameters(power=50e-6, waist=50e-6), laser_2=LaserParameters(power=float(p), waist=50e-6),
            Delta_e=2 * np.pi * 1e9))
    kw = dict(_plan_kwargs(inputs(p2[0])), laser_2_power=p2, tweezer_power=0.02, tweezer_waist=0.8e-6, warn=False)
    b = rg.simulate_CZ_gate_batch("lp", B=4, include_noise=True, phase_mode="compat", prologue="host", **kw)
    d = rg.simulate_CZ_gate_batch("lp", B=4, include_noise=True, prologue="host", **kw)
    assert np.abs(b.f11_population - d.f11_population).max() < 1e-12
    for i, p in enumerate(p2):
        one = rg.simulate_CZ_gate(inputs(p), tweezer_power=0.02, tweezer_waist=0.8e-6, include_noise=True,
                                  phase_mode="compat")
        assert abs(one.avg_fidelity - b.avg_fidelity[i]) < 1e-9
        assert abs(one.phase_info["cz_phase_fidelity"] - b.cz_phase_fidelity[i]) < 1e-9
