"""The shared coefficient table of the clustered path (csrc/rydcz_cluster.cuh, "the shared coefficient table"):
a (sub)step takes the records of the grid entry next to (h Vc, rho / Vc^3) and rescales its operator by
lambda = h Vc / theta_n.  CPU only (tests/hostsim + a NumPy restatement of the lookup rule)."""
import numpy as np
import pytest

import _cases as C
from noisyquantumsimulator_b200.codegen import model as M
from test_cluster_hostsim import _geometry, _pvec

REF = C.load("reference_cases.json")

CLASSES, PER_OCT, DTHETA, Z_MAX = 64, 8.0, 0.5, 27.0


def _lookup(h, Vc, w0, w1):
    """cluster_table_lookup restated: (theta_n, lambda, rho^_c) or None."""
    theta = abs(h) * Vc
    n = np.rint(theta / DTHETA)
    if n < 1:
        return None
    theta_n = n * DTHETA
    lam = theta / theta_n
    w1p = abs(lam - 1) + lam * w1 / Vc
    need = max(lam * w0 / Vc, (1 + w1p) * w1p * (2 + w1p))
    if not 0 < need <= 0.25:
        return None
    c = min(int(np.floor(PER_OCT * np.log2(0.25 / need))), CLASSES - 1)
    rho = 0.25 * 2.0 ** (-c / PER_OCT)
    if rho < need:
        c -= 1
        rho = 0.25 * 2.0 ** (-c / PER_OCT)
    return (theta_n, lam, rho) if c >= 0 else None


@pytest.mark.parametrize("dim,sym", [(3, True), (3, False), (4, True)])
def test_rescaled_spectrum_lies_inside_the_table_entrys_clusters(dim, sym):
    """What makes the shared records exact: every eigenvalue x of lambda G / Vc satisfies |q^(x)| <= rho^_c for the
    entry the lookup picks (q^(x) = x (x^2 + 1)), up to the damping the expansion tolerates anyway."""
    rng = np.random.default_rng(17 + dim + sym)
    systems = M.symmetric_systems(dim) if sym else M.standard_systems(dim)
    names = ["rho11s", "coh11s"] if sym else ["rho11", "coh11"]
    found = 0
    for _ in range(60):
        om, phi = 1.0, rng.uniform(-np.pi, np.pi)
        V = 10 ** rng.uniform(1.3, 3.0)
        det = rng.choice([0.0, rng.uniform(-1.0, 1.0)])
        detB = det if sym else det + 0.2 * rng.standard_normal()
        d1 = rng.choice([0.0, 0.06])
        r = 10 ** rng.uniform(-5, -2, size=5)
        rates = r[:4].sum() + (r[4] if dim == 4 else 0.0)
        Vc, w0, w1 = _geometry(om, det, detB, V, d1, rates, sym)
        w0, w1 = 1.02 * w0, 1.02 * w1          # (cluster_geometry's slack)
        h = rng.uniform(0.05, 6.0) * rng.choice([1.0, -1.0])
        hit = _lookup(h, Vc, w0, w1)
        if hit is None:
            continue
        found += 1
        theta_n, lam, rho = hit
        for nm in names:
            s = systems[nm]
            ev = np.linalg.eigvals(s.generator(_pvec(s, om, phi, det, detB, V, d1, r)[:len(s.M)])) * (lam / Vc)
            y = ev.imag                                     # x = i y (+ a small real part from the damping)
            q = np.abs(y * (1 - y * y))
            assert q.max() <= rho * (1 + 1e-9), (nm, V, h, q.max() / rho)
    assert found > 40


@pytest.fixture(scope="module")
def hostsim():
    return C.load_hostsim()


@pytest.mark.parametrize("case", [c for c in REF if c["dim"] in (3, 4)], ids=lambda c: c["name"])
def test_table_records_give_the_same_states_as_own_records(hostsim, case):
    noisy = case["rates"] is not None
    pts, seg, ns = C.pack_batch([case])
    a, sa = hostsim(case["dim"], noisy, pts, seg, ns, cluster=2, table=0)
    b, sb = hostsim(case["dim"], noisy, pts, seg, ns, cluster=2, table=1)
    assert np.abs(sa - sb).max() < 2e-12 and np.abs(a[:, :14] - b[:, :14]).max() < 2e-12
    gold = C.golden_finals(case)
    ops = C.states_to_operators(case["dim"], sb[0])
    for lab in ("01", "10", "11"):
        got = ops["rho" + lab] if noisy else ops["coh" + lab][:, 0]
        assert np.abs(got - gold[lab]).max() < 1e-11


def _population(n, seed, V_over_om=342.5):
    """n bang-bang candidates that share the apparatus (one V, one |Omega|): an optimiser generation."""
    rng = np.random.default_rng(seed)
    Om = 2 * np.pi * 4.6e6
    P = np.zeros((n, 12))
    P[:, 0] = Om * V_over_om
    P[:, 1] = 2 * np.pi * 1.2e5
    P[:, 4], P[:, 5], P[:, 6], P[:, 7] = 3.2e3, 4.1e3, 2.6e4, 8.9e3
    S = 5
    ot = rng.uniform(5, 40, n)
    b = np.concatenate([np.zeros((n, 1)), np.sort(rng.uniform(0.01, 0.99, (n, S - 1)), 1), np.ones((n, 1))], 1)
    ph = rng.uniform(-np.pi, np.pi, (n, S))
    G = np.zeros((n, S, 4))
    G[:, :, 0] = np.diff(b, axis=1) * (ot / Om)[:, None]
    G[:, :, 1], G[:, :, 2] = Om * np.cos(ph), Om * np.sin(ph)
    return P, G, np.full(n, S, dtype=np.int32)


@pytest.mark.parametrize("dim", [3, 4])
def test_a_population_shares_its_entries(hostsim, dim):
    """400 candidates x 5 slots: far fewer table entries than slots, nothing generated again by a second call in the
    persistent mode, everything generated again in the per-call mode; results equal to own records and to the plain
    path; the price is a few per cent more Clenshaw steps."""
    P, G, ns = _population(400, 3)
    plain, _ = hostsim(dim, True, P, G, ns, want_states=False, cluster=0)
    hostsim.table_generated()
    own, _ = hostsim(dim, True, P, G, ns, want_states=False, cluster=2, table=0)
    assert hostsim.table_generated() == 0
    tab, _ = hostsim(dim, True, P, G, ns, want_states=False, cluster=2, table=1)
    g1 = hostsim.table_generated()
    hostsim(dim, True, P, G, ns, want_states=False, cluster=2, table=1)
    assert hostsim.table_generated() == g1                       # per call: generated again
    hostsim(dim, True, P, G, ns, want_states=False, cluster=2, table=2)
    hostsim.table_generated()
    keep, _ = hostsim(dim, True, P, G, ns, want_states=False, cluster=2, table=2)
    assert hostsim.table_generated() == 0                        # kept across calls
    slots = 400 * 5
    assert 0 < g1 < slots                                        # entries are shared (more so in larger batches)
    assert np.abs(own[:, :14] - tab[:, :14]).max() < 2e-12 and np.array_equal(tab, keep)
    assert np.abs(plain[:, :14] - tab[:, :14]).max() < 2e-11
    st = C.OUT["steps"]
    assert tab[:, st].sum() < 1.06 * own[:, st].sum()


def test_slots_outside_the_table_keep_their_own_records(hostsim):
    """theta = h Vc < 0.25 (a sliver of a segment) has no grid entry: that slot generates its records itself while
    the others of the same point use the table."""
    P, G, ns = _population(8, 5)
    G[:, 2, 0] = 1e-4 / P[0, 0]                     # h Vc = 1e-4
    own, so = hostsim(3, True, P, G, ns, cluster=2, table=0)
    hostsim.table_generated()
    tab, st = hostsim(3, True, P, G, ns, cluster=2, table=1)
    used, back = hostsim.cluster_stats()
    assert back == 0 and hostsim.table_generated() > 0
    assert np.abs(so - st).max() < 2e-12
