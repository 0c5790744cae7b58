"""The clustered-spectrum path (csrc/rydcz_cluster.cuh) on the GPU: same results as the plain Chebyshev
path and the golden fixtures, the planner's choice, the hand-back of a point whose invariant check fails."""
import numpy as np
import pytest

import _cases as C

pytestmark = pytest.mark.gpu

REF = C.load("reference_cases.json")


@pytest.fixture()
def ev(evaluator):
    yield evaluator
    evaluator.set_cluster(1)


@pytest.mark.parametrize("dim,noisy", [(3, True), (3, False), (4, True), (4, False)])
def test_forced_cluster_path_matches_plain_path_and_golden(ev, dim, noisy):
    cases = [c for c in REF if c["dim"] == dim and (c["rates"] is not None) == noisy]
    pts, seg, ns = C.pack_batch(cases)
    ev.set_cluster(0)
    a, sa = ev.evaluate_host(dim, noisy, pts, seg, ns, want_states=True)
    ev.cluster_stats()
    ev.set_cluster(2)
    b, sb = ev.evaluate_host(dim, noisy, pts, seg, ns, want_states=True)
    planned, back = ev.cluster_stats()
    closed = dim == 3 and not noisy      # every system has a closed form (rydcz_closed.cuh): nothing to plan, no Chebyshev step
    assert (planned == 0 if closed else planned >= len(cases) // 2) and back == 0
    assert np.abs(a[:, :14] - b[:, :14]).max() < 5e-11
    assert np.abs(sa - sb).max() < 5e-11
    # fewer operator applications where the path was taken
    assert (a[:, 14].sum() == 0 and b[:, 14].sum() == 0) if closed else b[:, 14].sum() < 0.7 * a[:, 14].sum()
    for i, c in enumerate(cases):
        gold = C.golden_finals(c)
        ops = C.states_to_operators(dim, sb[i])
        for lab in ("01", "10", "11"):
            got = ops["rho" + lab] if noisy else ops["coh" + lab][:, 0]
            assert np.abs(got - gold[lab]).max() < 1e-10, c["name"]


def test_planner_prefers_cluster_only_for_strong_blockade(ev):
    names = {"lp_noisy": True, "bb5_noisy": True, "lp_noisy_spacing6p5_vlt10": False, "lp_noisy_n50": False}
    for name, want in names.items():
        case = next(c for c in REF if c["name"] == name)
        pts, seg, ns = C.pack_batch([case])
        ev.set_cluster(1)
        ev.cluster_stats()
        ev.evaluate_host(3, True, pts, seg, ns)
        planned, _ = ev.cluster_stats()
        assert (planned == 1) == want, name


def test_mixed_batch_kinds(ev):
    """General (unequal atoms), plain folded and clustered points in one batch, unsorted and sorted."""
    rng = np.random.default_rng(11)
    cases = [c for c in REF if c["dim"] == 3 and c["rates"] is not None and len(c["segments"]["dt"]) <= 7]
    pts, seg, ns = C.pack_batch(cases)
    reps = 40
    pts = np.tile(pts, (reps, 1)); seg = np.tile(seg, (reps, 1, 1)); ns = np.tile(ns, reps)
    odd = rng.random(len(pts)) < 0.3
    pts[odd, 2] += 2 * np.pi * 1e4 * rng.standard_normal(odd.sum())      # PT_OFF_A != PT_OFF_B -> general kernel
    ev.set_cluster(0)
    a, _ = ev.evaluate_host(3, True, pts, seg, ns)
    ev.set_cluster(1)
    ev.cluster_stats()
    b, _ = ev.evaluate_host(3, True, pts, seg, ns)
    planned, back = ev.cluster_stats()
    assert 0 < planned < len(pts) and back == 0
    assert np.abs(a[:, :14] - b[:, :14]).max() < 5e-11
    ev.set_sort(False)
    try:
        c, _ = ev.evaluate_host(3, True, pts, seg, ns)
    finally:
        ev.set_sort(True)
    assert np.array_equal(b[:, :14], c[:, :14])


def test_repeated_segment_path_for_smooth_jp(ev):
    """Smooth-JP fixtures (300 steps that differ only in the laser phase) take the one-propagator path by
    default; same states as the plain path and the golden finals."""
    for noisy in (True, False):
        cases = [c for c in REF if c["dim"] == 3 and (c["rates"] is not None) == noisy and len(c["segments"]["dt"]) == 300]
        assert len(cases) >= 2
        pts, seg, ns = C.pack_batch(cases)
        reps = 70                                   # > 64 points: the sorted order is used
        P, S, N = np.tile(pts, (reps, 1)), np.tile(seg, (reps, 1, 1)), np.tile(ns, reps)
        ev.set_cluster(0)
        a, sa = ev.evaluate_host(3, noisy, P, S, N, want_states=True)
        ev.repeat_stats()
        ev.set_cluster(1)
        b, sb = ev.evaluate_host(3, noisy, P, S, N, want_states=True)
        # (noiseless identical atoms: closed forms only, rydcz_closed.cuh -- no planned path, no Chebyshev step)
        assert ev.repeat_stats() == (len(P) if noisy else 0)
        assert np.abs(a[:, :14] - b[:, :14]).max() < 5e-11 and np.abs(sa - sb).max() < 5e-11
        assert b[:, 14].sum() < 0.2 * a[:, 14].sum() if noisy else (a[:, 14].sum() == 0 and b[:, 14].sum() == 0)
        b1, sb1 = ev.evaluate_host(3, noisy, pts, seg, ns, want_states=True)     # unsorted small batch
        assert np.array_equal(b1[:, :14], b[:len(cases), :14])
        for i, c in enumerate(cases):
            gold = C.golden_finals(c)
            ops = C.states_to_operators(3, sb[i])
            for lab in ("01", "10", "11"):
                got = ops["rho" + lab] if noisy else ops["coh" + lab][:, 0]
                assert np.abs(got - gold[lab]).max() < 1e-10, c["name"]


def test_shots_planned_per_shot_match_the_in_kernel_sampling(ev):
    """Monte-Carlo shots: with the clustered path on, every shot is planned as a point of its own (perturbed row,
    general clustered kernel); same Philox draws as the plain path's in-kernel sampling, same per-shot rows."""
    case = next(c for c in REF if c["name"] == "lp_noisy")
    pts, seg, ns = C.pack_batch([case])
    rng = np.random.default_rng(3)
    B, shots = 96, 16
    P, S, N = np.tile(pts, (B, 1)), np.tile(seg, (B, 1, 1)), np.tile(ns, B)
    P[:, 0] *= 10 ** rng.uniform(-0.3, 0.3, B)
    P[:, 9] = 2 * np.pi * rng.uniform(2e4, 8e4, B)       # PT_SIG_DET
    P[:, 10] = rng.uniform(0.002, 0.01, B)               # PT_SIG_RELR
    ev.set_cluster(0)
    m0, s0 = ev.evaluate_shots_host(3, True, shots, 99, P, S, N, want_shots=True)
    ev.cluster_stats()
    ev.set_cluster(1)
    m1, s1 = ev.evaluate_shots_host(3, True, shots, 99, P, S, N, want_shots=True)
    planned, back = ev.cluster_stats()
    assert planned == B * shots and back == 0
    assert np.abs(s0[..., :14] - s1[..., :14]).max() < 5e-11
    assert np.abs(m0[:, :14] - m1[:, :14]).max() < 5e-11
    assert s1[..., 14].sum() < 0.2 * s0[..., 14].sum()
