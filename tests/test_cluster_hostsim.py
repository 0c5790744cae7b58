"""The clustered-spectrum path (csrc/rydcz_cluster.cuh) compiled for the host: spectral enclosure, coefficient
records, parity with the golden fixtures and the plain Chebyshev path.  CPU only (tests/hostsim)."""
import ctypes as Ct

import numpy as np
import pytest
from scipy.linalg import expm

import _cases as C
from noisyquantumsimulator_b200.codegen import model as M

REF = C.load("reference_cases.json")
DP = Ct.POINTER(Ct.c_double)


@pytest.fixture(scope="module")
def hostsim():
    return C.load_hostsim()


def _geometry(om, detA, detB, V, d1, rates, sym):
    """cluster_geometry (csrc/rydcz_cluster.cuh) restated: centre Vc and half-widths of the three clusters."""
    def lam(det):
        a, b = 0.5 * (d1 - det), 0.5 * np.sqrt((d1 + det) ** 2 + om * om)
        return a - b, a + b
    lmA, lpA = lam(detA)
    lmB, lpB = lam(detB)
    if sym:
        a, b = 0.5 * (3 * d1 - detA), 0.5 * np.sqrt((d1 + detA) ** 2 + 2 * om * om)
        m, Mx = min(0.0, lmA, a - b), max(0.0, lpA, a + b)
    else:
        m, Mx = min(0.0, lmA) + min(0.0, lmB), max(0.0, lpA) + max(0.0, lpB)
    Err = V - detA - detB
    eta = 0.5 * om * om / (Err - Mx)
    if sym:
        m -= eta
    W, d = Mx - m, 4.0 * rates
    return Err + 0.5 * (eta - Mx - m), W + d, 0.5 * (W + eta) + d


def _pvec(s, om, phi, detA, detB, V, d1, r):
    d = dict(om_re=om * np.cos(phi), om_im=om * np.sin(phi), det=detA, det_a=detA, det_b=detB, V=V, delta1=d1,
             g_r1=r[0], g_r0=r[1], g_phi=r[2], g_sc=r[3], g_mj=r[4], shift=0.0)
    return [d.get(k, 0.0) for k in s.params] + ([0.0] if len(s.M) > len(s.params) else [])


@pytest.mark.parametrize("dim,sym", [(3, True), (3, False), (4, True), (4, False)])
def test_every_eigenvalue_lies_in_one_of_the_three_clusters(dim, sym):
    """The enclosure the path rests on: Im(lambda) within w0 of 0 or within w1 of +-Vc, -damp <= Re(lambda) <= 0,
    for the density-matrix systems and for the (unshifted) coherence systems."""
    rng = np.random.default_rng(5 + dim + sym)
    systems = M.symmetric_systems(dim) if sym else M.standard_systems(dim)
    names = ["rho11s", "coh11s"] if sym else ["rho11", "coh11"]
    for _ in range(40):
        om, phi = 1.0, rng.uniform(-np.pi, np.pi)
        V = 10 ** rng.uniform(1.0, 3.0)
        det = rng.choice([0.0, rng.uniform(-1.5, 1.5)])
        detB = det if sym else det + 0.3 * rng.standard_normal()
        d1 = rng.choice([0.0, 0.06, rng.uniform(-0.5, 0.5)])
        r = 10 ** rng.uniform(-5, -1.5, size=5)
        rates = r[:4].sum() + (r[4] if dim == 4 else 0.0)
        Vc, w0, w1 = _geometry(om, det, detB, V, d1, rates, sym)
        for nm in names:
            s = systems[nm]
            ev = np.linalg.eigvals(s.generator(_pvec(s, om, phi, det, detB, V, d1, r)[:len(s.M)]))
            dist0, dist1 = np.abs(ev.imag), np.abs(np.abs(ev.imag) - Vc)
            assert np.all((dist0 <= w0 * (1 + 1e-9)) | (dist1 <= w1 * (1 + 1e-9))), (nm, V, det, d1)
            assert ev.real.max() < 1e-9 and ev.real.min() > -4.0 * rates - 1e-9


def _records(hostsim, dim, noisy, pt, seg, ns):
    lib = hostsim.lib
    lib.hostsim_cluster_records.argtypes = [Ct.c_int, Ct.c_int, DP, DP, Ct.c_int, DP, DP]
    rec = np.zeros(8 * 65 * 4)
    cost = Ct.c_double(0)
    n = lib.hostsim_cluster_records(dim, int(noisy), np.ascontiguousarray(pt).ctypes.data_as(DP),
                                    np.ascontiguousarray(seg).ctypes.data_as(DP), int(ns), rec.ctypes.data_as(DP),
                                    Ct.byref(cost))
    return n, rec.reshape(-1, 4), cost.value


def test_coefficient_records_reproduce_the_matrix_exponential(hostsim):
    """exp(h G) x assembled in NumPy from the records of one LP pulse equals scipy's expm."""
    case = next(c for c in REF if c["name"] == "lp_noisy")
    pts, seg, ns = C.pack_batch([case])
    n, rec, cost = _records(hostsim, 3, True, pts[0], seg[0], ns[0])
    assert n > 0 and cost > 0
    K, nsub = int(rec[0, 0]) & 1023, int(rec[0, 0]) >> 10
    kappa, inv_gs, count = rec[0, 1], rec[0, 2], int(rec[0, 3])
    # both LP pulses share one slot; region = 8 slot headers (16 records), then the K + 1 terms of that slot
    assert count == 2 and nsub == 1 and n == 16 + K + 1 and rec[1, 0] == 16
    d = rec[16:16 + K + 1, :3]
    assert np.abs(d[-2:]).max() < 1e-13 * np.abs(d).max()     # the series has converged
    s = M.symmetric_systems(3)["rho11s"]
    row, sg = pts[0], seg[0, 0]
    p = dict(om_re=sg[1], om_im=sg[2], det=sg[3], V=row[0], delta1=row[1], g_r1=row[4], g_r0=row[5], g_phi=row[6],
             g_sc=row[7])
    G = s.generator([p[k] for k in s.params])
    Gt = G * inv_gs
    S2 = Gt @ (Gt @ Gt + kappa * np.eye(s.n))                 # 2 S~
    x = s.x0.copy()
    phi_prev, phi = x.copy(), 0.5 * S2 @ x
    Cx = lambda k, v: d[k, 0] * v + d[k, 1] * (Gt @ v) + d[k, 2] * (Gt @ (Gt @ v))
    y = Cx(0, phi_prev) + Cx(1, phi)
    for k in range(2, K + 1):
        phi, phi_prev = S2 @ phi + phi_prev, phi
        y = y + Cx(k, phi)
    assert np.abs(y - expm(sg[0] * G) @ x).max() < 5e-13


@pytest.mark.parametrize("case", [c for c in REF if c["dim"] in (3, 4)], ids=lambda c: c["name"])
def test_states_and_fidelity_match_golden_on_the_clustered_path(hostsim, case):
    noisy = case["rates"] is not None
    pts, seg, ns = C.pack_batch([case])
    out, st = hostsim(case["dim"], noisy, pts, seg, ns, cluster=2)
    used, handed_back = hostsim.cluster_stats()
    assert handed_back == 0
    gold = C.golden_finals(case)
    ops = C.states_to_operators(case["dim"], st[0])
    for lab in ("01", "10", "11"):
        got = ops["rho" + lab] if noisy else ops["coh" + lab][:, 0]
        assert np.abs(got - gold[lab]).max() < 1e-11
    res = case["result_coherence"] if noisy else case["result"]
    assert abs(out[0, C.OUT["avg"]] - res["avg_fidelity"]) < 1e-10
    assert abs(out[0, C.OUT["czf"]] - res["phase_info"]["cz_phase_fidelity"]) < 1e-10


def test_planner_choice_and_work_reduction(hostsim):
    """cluster=1 takes the path only where it is cheaper: strong blockade yes, V/Omega < 10 and 998 distinct
    envelope segments no; a bang-bang candidate then needs > 10x fewer operator applications."""
    want = {"bb5_noisy": True, "lp_noisy": True, "lp_noisy_spacing6p5_vlt10": False, "lp_shaped_cosine_noisy": False}
    for name, expect in want.items():
        case = next(c for c in REF if c["name"] == name)
        pts, seg, ns = C.pack_batch([case])
        a, _ = hostsim(3, True, pts, seg, ns, cluster=0)
        hostsim.cluster_stats()
        b, _ = hostsim(3, True, pts, seg, ns, cluster=1)
        used, _ = hostsim.cluster_stats()
        assert (used == 1) == expect, name
        assert np.abs(a[0, :14] - b[0, :14]).max() < 1e-11
        if name == "bb5_noisy":
            assert b[0, C.OUT["steps"]] < 0.1 * a[0, C.OUT["steps"]]


def test_strong_dephasing_and_negative_duration_take_substeps(hostsim):
    """gamma tau > 1 is cut into sub-steps (damp * h <= 1); dt < 0 (the reference's Delta_e < 0 quirk) runs
    the expansion backwards.  Both against the plain path."""
    for name in ("lp_noisy_linewidth_1MHz", "lp_noisy_neg_Delta_e"):
        case = next(c for c in REF if c["name"] == name)
        pts, seg, ns = C.pack_batch([case])
        a, sa = hostsim(3, True, pts, seg, ns, cluster=0)
        hostsim.cluster_stats(); hostsim.repeat_used()
        b, sb = hostsim(3, True, pts, seg, ns, cluster=2)
        used, back = hostsim.cluster_stats()
        # (the 300 equal steps of the negative-Delta_e smooth-JP case go to the repeated-segment path instead)
        assert used + hostsim.repeat_used() == 1 and back == 0
        assert np.abs(sa - sb).max() < 1e-11
    n, rec, _ = _records(hostsim, 3, True, *[x[0] for x in C.pack_batch(
        [next(c for c in REF if c["name"] == "lp_noisy_linewidth_1MHz")])])
    assert (int(rec[0, 0]) >> 10) > 1


def test_unequal_atoms_take_the_general_systems_on_the_clustered_path(hostsim):
    """Doppler-shifted atoms (off_a != off_b, as in a Monte-Carlo shot) and a perturbed blockade: the general
    25-variable rho_11 and 8-variable coherence on the clustered expansion against the plain path."""
    rng = np.random.default_rng(7)
    for name in ("lp_noisy", "bb5_noisy", "bb7_noisy", "vb_bb5_default", "lp_noisy_cs133"):
        case = next(c for c in REF if c["name"] == name)
        noisy = case["rates"] is not None
        pts, seg, ns = C.pack_batch([case])
        pts, seg, ns = np.tile(pts, (4, 1)), np.tile(seg, (4, 1, 1)), np.tile(ns, 4)
        pts[:, 2] += 2 * np.pi * 3e4 * rng.standard_normal(4)          # PT_OFF_A
        pts[:, 3] += 2 * np.pi * 3e4 * rng.standard_normal(4)          # PT_OFF_B
        pts[:, 0] *= 1 + 0.01 * rng.standard_normal(4)                 # PT_V
        a, sa = hostsim(3, noisy, pts, seg, ns, cluster=0)
        hostsim.cluster_stats()
        b, sb = hostsim(3, noisy, pts, seg, ns, cluster=2)
        used, back = hostsim.cluster_stats()
        # (noiseless: every system has a closed form, rydcz_closed.cuh -- nothing is planned, no Chebyshev step)
        assert used == (4 if noisy else 0) and back == 0, name
        assert np.abs(sa - sb).max() < 1e-11 and np.abs(a[:, :14] - b[:, :14]).max() < 1e-11, name
        if noisy:
            assert b[:, C.OUT["steps"]].sum() < 0.2 * a[:, C.OUT["steps"]].sum()
        else:
            assert b[:, C.OUT["steps"]].sum() == 0 and a[:, C.OUT["steps"]].sum() == 0


@pytest.mark.parametrize("dim,noisy", [(3, True), (3, False), (4, True)])
def test_random_points_agree_on_both_paths(hostsim, dim, noisy):
    """Seeded random LP and bang-bang points in the SURVEY 8d ranges (V/Omega 4-800, identical and, for the
    3-level model, unequal atoms): the planner's choice (cluster=1) against the single-interval path."""
    rng = np.random.default_rng(100 * dim + noisy)
    n = 60
    Om = 2 * np.pi * rng.uniform(1.5e6, 9e6, n)
    P = np.zeros((n, 12))
    P[:, 0] = Om * 10 ** rng.uniform(0.6, 2.9, n)
    P[:, 1] = 2 * np.pi * rng.uniform(0, 3e5, n)
    P[:, 4], P[:, 5] = rng.uniform(2e3, 5e3, n), rng.uniform(2e3, 9e3, n)
    P[:, 6], P[:, 7], P[:, 8] = 10 ** rng.uniform(2, 6, n), 10 ** rng.uniform(3, 5.5, n), rng.uniform(0, 5e3, n)
    if dim == 3:
        odd = rng.random(n) < 0.4
        P[odd, 2] = 2 * np.pi * 5e4 * rng.standard_normal(odd.sum())
        P[odd, 3] = 2 * np.pi * 5e4 * rng.standard_normal(odd.sum())
    S = 5
    G = np.zeros((n, S, 4))
    ns = np.full(n, S, dtype=np.int32)
    lp = rng.random(n) < 0.4
    ns[lp] = 2
    tau = rng.uniform(3.5, 5.5, n) / Om
    xi = np.exp(1j * rng.uniform(-np.pi, np.pi, n))
    ot = rng.uniform(5, 40, n)
    b = np.concatenate([np.zeros((n, 1)), np.sort(rng.uniform(0.01, 0.99, (n, S - 1)), 1), np.ones((n, 1))], 1)
    ph = rng.uniform(-np.pi, np.pi, (n, S))
    G[:, :, 0] = np.diff(b, axis=1) * (ot / Om)[:, None]
    G[:, :, 1], G[:, :, 2] = Om[:, None] * np.cos(ph), Om[:, None] * np.sin(ph)
    G[lp, 0, 0] = G[lp, 1, 0] = tau[lp]
    G[lp, 0, 1], G[lp, 0, 2] = Om[lp], 0.0
    G[lp, 1, 1], G[lp, 1, 2] = (Om * xi).real[lp], (Om * xi).imag[lp]
    G[lp, :2, 3] = (rng.uniform(0.2, 0.5, n) * Om)[lp, None]
    a, sa = hostsim(dim, noisy, P, G, ns, cluster=0)
    hostsim.cluster_stats()
    b_, sb = hostsim(dim, noisy, P, G, ns, cluster=1)
    used, back = hostsim.cluster_stats()
    # weak-blockade points stay on the plain path -- and so do noiseless identical atoms of the 3-level model, whose
    # systems all have closed forms (rydcz_closed.cuh): nothing is planned there
    assert (used == 0 if (dim == 3 and not noisy) else n // 3 < used < n) and back == 0
    assert np.abs(sa - sb).max() < 1e-11
    assert np.abs(a[:, :14] - b_[:, :14]).max() < 5e-11
