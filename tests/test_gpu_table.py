"""The shared coefficient table of the clustered path on the GPU (csrc/rydcz_cluster.cuh; rydcz_set_table): the three
modes agree with each other and with the plain path, a population shares its entries, the chunks of a pipelined
call share them across streams."""
import numpy as np
import pytest

import _cases as C
from test_table_hostsim import _population

pytestmark = pytest.mark.gpu

REF = C.load("reference_cases.json")


@pytest.fixture()
def ev(evaluator):
    yield evaluator
    evaluator.set_cluster(1)
    evaluator.set_table(1)


# (3-level noiseless identical atoms are closed forms, rydcz_closed.cuh: no clustered slots; the 4-level model covers that branch)
@pytest.mark.parametrize("dim,noisy", [(3, True), (4, False), (4, True)])
def test_population_shares_table_entries(ev, dim, noisy):
    n = 20000
    P, G, ns = _population(n, 7)
    if not noisy:
        P[:, 4:9] = 0.0
    ev.set_cluster(0)
    plain, sp = ev.evaluate_host(dim, noisy, P, G, ns, want_states=True)
    ev.set_cluster(2)
    ev.cluster_stats(); ev.table_stats()
    ev.set_table(0)
    own, so = ev.evaluate_host(dim, noisy, P, G, ns, want_states=True)
    assert ev.table_stats() == (0, 0)
    ev.set_table(1)
    tab, st = ev.evaluate_host(dim, noisy, P, G, ns, want_states=True)
    gen1, shared1 = ev.table_stats()
    tab2, _ = ev.evaluate_host(dim, noisy, P, G, ns, want_states=True)
    gen2, shared2 = ev.table_stats()
    ev.set_table(2)
    ev.evaluate_host(dim, noisy, P, G, ns)
    ev.table_stats()
    keep, sk = ev.evaluate_host(dim, noisy, P, G, ns, want_states=True)
    gen3, shared3 = ev.table_stats()
    planned, back = ev.cluster_stats()
    assert back == 0 and planned == 5 * n
    # every slot of every candidate reads the table; 100,000 slots map onto a few thousand entries; the per-call mode
    # generates them again, the persistent mode does not
    assert shared1 == shared2 == shared3 and 0.99 * 5 * n < shared1 <= 5 * n     # (a sliver of a segment has no entry)
    assert 0 < gen1 < 5 * n // 4 and gen2 == gen1 and gen3 == 0      # (at most ~31,000 entries for any batch size)
    assert np.array_equal(tab, tab2) and np.array_equal(tab, keep) and np.array_equal(st, sk)
    assert np.abs(own[:, :14] - tab[:, :14]).max() < 5e-12 and np.abs(so - st).max() < 5e-12
    assert np.abs(plain[:, :14] - tab[:, :14]).max() < 5e-11 and np.abs(sp - st).max() < 5e-11
    assert tab[:, 14].sum() < 1.06 * own[:, 14].sum()            # Clenshaw steps: a few per cent more


def test_fixtures_on_the_table(ev):
    """Every 3- and 4-level fixture through table records (forced clustered path) against the golden final states."""
    for dim in (3, 4):
        for noisy in (True, False):
            cases = [c for c in REF if c["dim"] == dim and (c["rates"] is not None) == noisy]
            if not cases:
                continue
            pts, seg, ns = C.pack_batch(cases)
            ev.set_cluster(2)
            ev.set_table(1)
            ev.table_stats()
            out, st = ev.evaluate_host(dim, noisy, pts, seg, ns, want_states=True)
            gen, shared = ev.table_stats()
            if dim == 3 and not noisy:          # closed forms only: nothing is planned
                assert gen == 0 and shared == 0
            else:
                assert gen > 0 and shared >= gen
            for i, c in enumerate(cases):
                gold = C.golden_finals(c)
                ops = C.states_to_operators(dim, st[i])
                for lab in ("01", "10", "11"):
                    got = ops["rho" + lab] if noisy else ops["coh" + lab][:, 0]
                    assert np.abs(got - gold[lab]).max() < 1e-10, c["name"]


def test_unequal_atoms_and_mixed_kinds_on_the_table(ev):
    """Doppler-shifted atoms (general clustered kernel), identical atoms and weak-blockade points in one batch."""
    rng = np.random.default_rng(21)
    n = 6000
    P, G, ns = _population(n, 9)
    odd = rng.random(n) < 0.5
    P[odd, 2] = 2 * np.pi * 4e4 * rng.standard_normal(odd.sum())
    P[odd, 3] = 2 * np.pi * 4e4 * rng.standard_normal(odd.sum())
    P[:, 0] *= 10 ** rng.uniform(-1.8, 0.3, n)                   # V / Omega from 5 to 680
    ev.set_cluster(0)
    plain, sp = ev.evaluate_host(3, True, P, G, ns, want_states=True)
    ev.set_cluster(1)
    ev.cluster_stats(); ev.table_stats()
    tab, st = ev.evaluate_host(3, True, P, G, ns, want_states=True)
    planned, back = ev.cluster_stats()
    gen, shared = ev.table_stats()
    assert planned > n // 2 and back == 0 and gen > 0 and shared > planned
    assert np.abs(plain[:, :14] - tab[:, :14]).max() < 5e-11 and np.abs(sp - st).max() < 5e-11


def test_pipelined_chunks_share_the_table_across_streams(ev):
    """rydcz_simulate_host_cols cuts a batch into chunks on eight streams; a chunk may rely on an entry that another
    chunk's planner claimed.  Same results as one chunk, and as the table switched off."""
    from noisyquantumsimulator_b200 import rydberg_gates as rg
    from noisyquantumsimulator_b200.engine import get_evaluator
    import bench
    n = 50000
    cand = bench.make_candidates(n, 5)
    d = get_evaluator(0)
    try:
        d.set_table(0)
        a = rg.simulate_CZ_gate_batch("jp", B=n, warn=False, **bench.MEDIUM, **cand)
        d.set_table(1)
        d.table_stats()
        d.set_chunk(4096)
        b = rg.simulate_CZ_gate_batch("jp", B=n, warn=False, **bench.MEDIUM, **cand)
        gen, shared = d.table_stats()
        d.set_chunk(1 << 20)
        c = rg.simulate_CZ_gate_batch("jp", B=n, warn=False, **bench.MEDIUM, **cand)
    finally:
        d.set_chunk(32768)
        d.set_table(1)
    assert 0.99 * 5 * n < shared <= 5 * n and 0 < gen < n
    assert np.array_equal(b.raw[:, :14], c.raw[:, :14])
    assert np.abs(a.raw[:, :14] - b.raw[:, :14]).max() < 5e-12
