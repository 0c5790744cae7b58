"""The repeated-segment path (csrc/rydcz_repeat.cuh) compiled for the host: the laser-phase covariance it rests
on, parity with the golden fixtures and the plain path for the smooth-JP protocol.  CPU only (tests/hostsim)."""
import numpy as np
import pytest

import _cases as C
from noisyquantumsimulator_b200.codegen import model as M

REF = C.load("reference_cases.json")
SJP = [c for c in REF if c["dim"] == 3 and len(c["segments"]["dt"]) == 300]


@pytest.fixture(scope="module")
def hostsim():
    return C.load_hostsim()


def _rotation(s, phi):
    """R(phi) as codegen/emit.py emits it: the (re, im) pair of X[i,j] turned by q phi, q = Rydberg
    excitations of |i> minus those of |j>."""
    d = s.d
    nr = lambda i: sum(1 for lev in divmod(int(i), d) if lev >= 2)
    R = np.eye(s.n)
    for k, (i, j, part) in enumerate(s.var_info):
        if part != "re" or (i, j, "im") not in s.var_info:
            continue
        k2, a = s.var_info.index((i, j, "im")), (nr(i) - nr(j)) * phi
        R[k, k], R[k, k2], R[k2, k], R[k2, k2] = np.cos(a), -np.sin(a), np.sin(a), np.cos(a)
    return R


@pytest.mark.parametrize("dim", [3, 4])
def test_generator_is_covariant_under_the_laser_phase(dim):
    """G(Omega e^{i phi}) = R(phi) G(Omega) R(phi)^T for every reduced system, with all rates on."""
    rng = np.random.default_rng(3)
    systems = dict(M.standard_systems(dim))
    systems.update(M.symmetric_systems(dim))
    for name, s in systems.items():
        for _ in range(3):
            phi = rng.uniform(-np.pi, np.pi)
            vals = dict(det=0.3, det_a=0.3, det_b=-0.2, V=20.0, delta1=0.06, g_r1=1e-2, g_r0=2e-2, g_phi=3e-2,
                        g_sc=1e-2, g_mj=5e-3)
            p0 = [dict(vals, om_re=1.3, om_im=0.0)[k] for k in s.params]
            p1 = [dict(vals, om_re=1.3 * np.cos(phi), om_im=1.3 * np.sin(phi))[k] for k in s.params]
            R = _rotation(s, phi)
            assert np.abs(R @ s.generator(p0) @ R.T - s.generator(p1)).max() < 1e-13, name


@pytest.mark.parametrize("case", SJP, ids=lambda c: c["name"])
def test_smooth_jp_matches_golden_on_the_repeated_segment_path(hostsim, case):
    noisy = case["rates"] is not None
    pts, seg, ns = C.pack_batch([case])
    a, sa = hostsim(3, noisy, pts, seg, ns, cluster=0)
    hostsim.repeat_used()
    out, st = hostsim(3, noisy, pts, seg, ns, cluster=1)
    # (noiseless, identical atoms: every system is a coherence with a closed form, rydcz_closed.cuh -- no Chebyshev
    #  step at all, and the planner keeps the point off the planned paths)
    assert hostsim.repeat_used() == (1 if noisy else 0)
    assert np.abs(sa - st).max() < 1e-11 and np.abs(a[0, :14] - out[0, :14]).max() < 1e-11
    if noisy:
        # operator applications: > 6x fewer (the plain path counts no steps for rho_01 either: closed form)
        assert out[0, C.OUT["steps"]] < 0.16 * a[0, C.OUT["steps"]]
    else:
        assert out[0, C.OUT["steps"]] == 0 and a[0, C.OUT["steps"]] == 0
    gold = C.golden_finals(case)
    ops = C.states_to_operators(3, st[0])
    for lab in ("01", "10", "11"):
        got = ops["rho" + lab] if noisy else ops["coh" + lab][:, 0]
        assert np.abs(got - gold[lab]).max() < 1e-11
    res = case["result_coherence"] if noisy else case["result"]
    assert abs(out[0, C.OUT["avg"]] - res["avg_fidelity"]) < 1e-10


def test_not_taken_when_segments_differ(hostsim):
    """Bang-bang (5 different durations), shaped LP (998 different amplitudes), and dim 4 stay off the path."""
    for name in ("bb5_noisy", "lp_shaped_cosine_noisy", "lp_noisy"):
        case = next(c for c in REF if c["name"] == name)
        pts, seg, ns = C.pack_batch([case])
        hostsim.repeat_used()
        hostsim(case["dim"], True, pts, seg, ns, cluster=1)
        assert hostsim.repeat_used() == 0, name
