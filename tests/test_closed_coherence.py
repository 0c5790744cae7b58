"""Closed-form |ab><00| coherences of the 3-level model (csrc/rydcz_closed.cuh) against 40-digit matrix
exponentials of the code generator's own generators (codegen/model.py), through the host build of the header."""
import ctypes as Ct

import mpmath as mp
import numpy as np
import pytest

import _cases as C
from noisyquantumsimulator_b200.codegen import model as M

DP = Ct.POINTER(Ct.c_double)


@pytest.fixture(scope="module")
def hostsim():
    return C.load_hostsim()


def _closed(hostsim, which, noisy, pt, segs):
    lib = hostsim.lib
    lib.hostsim_closed_coherence.argtypes = [Ct.c_int, Ct.c_int, DP, DP, Ct.c_int, DP]
    x = np.zeros({0: 6, 3: 8, 4: 5, 5: 5}.get(which, 4))
    segs = np.ascontiguousarray(segs, dtype=np.float64)
    ok = lib.hostsim_closed_coherence(which, int(noisy), np.ascontiguousarray(pt).ctypes.data_as(DP),
                                      segs.ctypes.data_as(DP), len(segs), x.ctypes.data_as(DP))
    return ok, x


def _exact(system, pt, segs, noisy, det_key):
    """exp(h_S G_S) ... exp(h_1 G_1) x0 with 40 digits."""
    mp.mp.dps = 40
    x = mp.matrix(system.x0.tolist())
    for dt, re, im, det in segs:
        if dt == 0.0:
            continue
        vals = dict(om_re=re, om_im=im, det=det + pt[2], det_a=det + pt[2], det_b=det + pt[3], V=pt[0], delta1=pt[1],
                    g_r1=pt[4] * noisy, g_r0=pt[5] * noisy, g_phi=pt[6] * noisy, g_sc=pt[7] * noisy)
        G = system.generator([vals[k] for k in system.params])
        x = mp.expm(mp.matrix((dt * G).tolist())) * x
    return np.array([float(t) for t in x])


def _random_point(rng, trial):
    Om = 2 * np.pi * rng.uniform(1e6, 6e6)
    V = Om * 10 ** rng.uniform(-1.5, 3)
    det = Om * rng.uniform(-1, 1)
    d1 = 2 * np.pi * rng.uniform(-3e5, 3e5)
    if trial % 11 == 0:
        Om = 0.0
    if trial % 13 == 0:
        det = -d1                       # d0 = d1 on the diagonal
    if trial % 17 == 0:
        V = det + d1                    # d2 ~ d1
    if trial % 19 == 0:
        d1 = 0.0
    pt = np.zeros(12)
    pt[[0, 1, 4, 5, 6, 7]] = [V, d1, rng.uniform(0, 5e3), rng.uniform(0, 5e3), rng.uniform(0, 1e5), rng.uniform(0, 1e3)]
    n_seg = int(rng.integers(1, 6))
    segs = np.zeros((n_seg, 4))
    scale = max(Om, 2 * np.pi * 1e6)
    for s in range(n_seg):
        ph = rng.uniform(-np.pi, np.pi)
        segs[s] = [rng.uniform(0.05, 8) / scale * (1 if (trial + s) % 5 else -1), Om * np.cos(ph), Om * np.sin(ph), det]
    if trial % 4 == 0 and n_seg > 1:
        segs[1:, 0] = segs[0, 0]        # repeated durations: the cached exponentials
    if trial % 6 == 0 and n_seg > 2:
        segs[1, 0] = 0.0                # an inactive row
    return pt, segs


@pytest.mark.parametrize("noisy", [True, False])
def test_folded_11_coherence_against_40_digit_expm(hostsim, noisy):
    rng = np.random.default_rng(7 + noisy)
    system = M.symmetric_systems(3)["coh11s"]
    worst = 0.0
    for trial in range(120):
        pt, segs = _random_point(rng, trial)
        ok, x = _closed(hostsim, 0, noisy, pt, segs)
        y = _exact(system, pt, segs, noisy, "det")
        tol = 2e-13 + 4e-16 * np.abs(segs[:, 0] * pt[0]).sum()       # the phase h V itself carries eps |h V|
        err = np.abs(x - y).max() / max(np.abs(y).max(), 1e-300)
        worst = max(worst, err / tol)
        assert ok == 1 and err < tol, (trial, err, tol)
    assert worst < 1.0


@pytest.mark.parametrize("which", [1, 2])
def test_single_atom_coherence_against_40_digit_expm(hostsim, which):
    rng = np.random.default_rng(70 + which)
    system = M.standard_systems(3)["coh01" if which == 1 else "coh10"]
    for trial in range(80):
        pt, segs = _random_point(rng, trial)
        pt[2], pt[3] = 2 * np.pi * rng.normal(0, 5e4, 2)             # per-atom detuning offsets
        ok, x = _closed(hostsim, which, True, pt, segs)
        y = _exact(system, pt, segs, 1.0, "det_b" if which == 1 else "det_a")
        err = np.abs(x - y).max() / max(np.abs(y).max(), 1e-300)
        assert ok == 1 and err < 2e-13, (trial, err)


@pytest.mark.parametrize("which", [4, 5])
def test_single_atom_rho_against_40_digit_expm(hostsim, which):
    """rho_01 / rho_10: the optical Bloch equations with decay as a real 4 x 4 closed form (fast pair by Newton, slow
    pair from the deflated quadratic).  Accepted whenever drive or detuning dominate the rates -- every pulse segment
    of the path -- and accurate wherever it is accepted; an idle segment at zero detuning is declined."""
    rng = np.random.default_rng(40 + which)
    system = M.standard_systems(3)["rho01" if which == 4 else "rho10"]
    declined = 0
    for trial in range(120):
        pt, segs = _random_point(rng, trial)
        pt[2], pt[3] = 2 * np.pi * rng.normal(0, 5e4, 2)
        if trial % 7 == 0:
            pt[5] = 0.0                                     # no decay to |0>: a root at exactly 0
        if trial % 9 == 0:
            segs[:, 0] *= 25.0                              # long segments: |h Omega| up to 200
        ok, x = _closed(hostsim, which, True, pt, segs)
        om_eff = np.sqrt(segs[0, 1] ** 2 + segs[0, 2] ** 2 + (segs[0, 3] + pt[2 if which == 5 else 3] + pt[1]) ** 2)
        if not ok:
            declined += 1
            assert om_eff < 64 * (pt[4] + pt[5] + pt[6] + pt[7]), (trial, om_eff)
            continue
        y = _exact(system, pt, segs, 1.0, None)
        tol = 2e-14 + 4e-16 * np.abs(segs[:, 0]).sum() * om_eff
        err = np.abs(x - y).max()
        assert err < tol, (trial, err, tol)
        assert abs(x[0] + x[1] + x[4] - 1.0) < 1e-14
    assert declined < 12


def test_single_atom_rho_declines_what_it_cannot_separate(hostsim):
    """Strong damping against a weak drive: the closed form either declines or is accurate -- never wrong."""
    rng = np.random.default_rng(5)
    system = M.standard_systems(3)["rho01"]
    accepted = declined = 0
    for trial in range(150):
        Om = 10 ** rng.uniform(2, 8)
        pt = np.zeros(12)
        pt[4:8] = 10 ** rng.uniform(0, 6.5, 4)
        det = Om * rng.uniform(-3, 3) * (trial % 3 != 0)
        ph = rng.uniform(-np.pi, np.pi)
        segs = np.array([[rng.uniform(0.05, 8) / max(Om, 1e5), Om * np.cos(ph), Om * np.sin(ph), det]])
        ok, x = _closed(hostsim, 4, True, pt, segs)
        if not ok:
            declined += 1
            continue
        accepted += 1
        y = _exact(system, pt, segs, 1.0, None)
        assert np.abs(x - y).max() < 5e-14, (trial, np.abs(x - y).max())
    assert accepted > 30 and declined > 30


@pytest.mark.parametrize("noisy", [True, False])
def test_unequal_atoms_11_coherence_against_40_digit_expm(hostsim, noisy):
    """The general |11><00| coherence (4 x 4: the fast eigenvector split off, three slow roots by Newton on the
    quartic).  In the regime of a Doppler shot (V / Omega >= 16, offsets of tens of kHz) it is always accepted; with
    weak blockade or strongly unequal atoms it may decline (the kernels then run the Chebyshev loop) -- but whatever
    it accepts is accurate."""
    rng = np.random.default_rng(17 + noisy)
    system = M.standard_systems(3)["coh11"]
    declined = 0
    for trial in range(160):
        pt, segs = _random_point(rng, trial)
        realistic = trial % 2 == 0
        Om = max(np.hypot(segs[0, 1], segs[0, 2]), 2 * np.pi * 1e6)
        if realistic:
            pt[0] = Om * 10 ** rng.uniform(1.2, 2.9)
            pt[2], pt[3] = 2 * np.pi * rng.normal(0, 8e4, 2)
            if np.hypot(segs[0, 1], segs[0, 2]) == 0.0:
                continue                                    # (an idle segment between unequal atoms: declined by design)
        else:
            pt[2], pt[3] = 2 * np.pi * rng.normal(0, 1e6, 2)
        ok, x = _closed(hostsim, 3, noisy, pt, segs)
        if not ok:
            declined += 1
            assert not realistic, trial
            continue
        y = _exact(system, pt, segs, noisy, None)
        tol = 2e-13 + 4e-16 * np.abs(segs[:, 0] * pt[0]).sum()
        err = np.abs(x - y).max() / max(np.abs(y).max(), 1e-300)
        assert err < tol, (trial, err, tol)
    assert 0 < declined < 80


def test_closed_single_atom_rho_against_the_chebyshev_build_of_the_same_evaluator(hostsim):
    """Path against path: the whole evaluation (all paths, identical and unequal atoms, LP / bang-bang) with the
    closed-form rho_01 / rho_10 equals the build that keeps them on the Chebyshev loop (-DRYDCZ_CLOSED_RHO01=0),
    states and result columns, and it counts fewer Chebyshev steps."""
    cheb = C.load_hostsim("chebrho", ["-DRYDCZ_CLOSED_RHO01=0"])
    rng = np.random.default_rng(321)
    B, S = 96, 5
    P = np.zeros((B, 12))
    G = np.zeros((B, S, 4))
    ns = np.full(B, S, dtype=np.int32)
    for i in range(B):
        Om = 2 * np.pi * rng.uniform(1.5e6, 9e6)
        P[i, 0] = Om * 10 ** rng.uniform(0.8, 2.6)
        P[i, 1] = 2 * np.pi * rng.uniform(0, 3e5)
        if i % 3 == 0:
            P[i, 2:4] = 2 * np.pi * rng.normal(0, 6e4, 2)         # unequal atoms
        P[i, 4:8] = [rng.uniform(2e3, 5e3), rng.uniform(2e3, 6e3), rng.uniform(1e2, 1e5), rng.uniform(1e3, 2e4)]
        if i % 2:                                                  # bang-bang: five phases, Delta = 0
            ot = rng.uniform(5, 40)
            t = np.concatenate([[0.0], np.sort(rng.uniform(0.01, 0.99, 4)) * ot, [ot]]) / Om
            ph = rng.uniform(-np.pi, np.pi, 5)
            G[i] = np.stack([np.diff(t), Om * np.cos(ph), Om * np.sin(ph), np.zeros(5)], 1)
        else:                                                      # LP: two pulses with a phase jump, Delta != 0
            tau, det, xi = rng.uniform(3.5, 5.5) / Om, rng.uniform(0.2, 0.5) * Om, rng.uniform(-np.pi, np.pi)
            G[i, 0] = [tau, Om, 0.0, det]
            G[i, 1] = [tau, Om * np.cos(xi), Om * np.sin(xi), det]
            ns[i] = 2
    for cluster in (0, 1):
        a, sa = hostsim(3, True, P, G, ns, cluster=cluster)
        b, sb = cheb(3, True, P, G, ns, cluster=cluster)
        assert np.abs(sa - sb).max() < 2e-12
        assert np.abs(a[:, :14] - b[:, :14]).max() < 2e-12
        assert np.all(a[:, C.OUT["steps"]] < b[:, C.OUT["steps"]])


def test_noiseless_identical_atoms_need_no_chebyshev_step(hostsim):
    """Every system of a noiseless identical-atom evaluation is a coherence: the fixture is reproduced with zero
    Chebyshev steps, states included."""
    case = next(c for c in C.load("reference_cases.json") if c["name"] == "vb_bb7_paper")
    pts, seg, ns = C.pack_batch([case])
    out, st = hostsim(3, False, pts, seg, ns)
    assert out[0, C.OUT["steps"]] == 0
    gold = C.golden_finals(case)
    ops = C.states_to_operators(3, st[0])
    for lab in ("01", "10", "11"):
        assert np.abs(ops["coh" + lab][:, 0] - gold[lab]).max() < 1e-11
    assert abs(out[0, C.OUT["avg"]] - case["result"]["avg_fidelity"]) < 1e-10
