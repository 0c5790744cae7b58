"""N > 1 path on CPU: two gloo ranks shard a population, all-gather the scores, and reproduce the
single-rank result bit for bit.  The per-rank arithmetic comes from the host harness (no GPU here);
on the GPU box the same function runs with engine.Evaluator and NCCL (bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import _cases as C


def _population(n=48):
    case = next(c for c in C.load("reference_cases.json") if c["name"] == "bb5_noisy")
    pt, seg = C.pack_case(case)
    rng = np.random.default_rng(42)
    P = np.repeat(pt[None], n, 0)
    G = np.repeat(seg[None], n, 0)
    G[:, :, 0] *= rng.uniform(0.2, 1.5, (n, 5))
    ph = rng.uniform(-np.pi, np.pi, (n, 5))
    amp = np.hypot(seg[0, 1], seg[0, 2])
    G[:, :, 1], G[:, :, 2] = amp * np.cos(ph), amp * np.sin(ph)
    return P, G


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import _cases as C2
    from noisyquantumsimulator_b200 import sharding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, G = _population()
    out = sharding.evaluate_sharded(C2.HostsimEvaluator(), 3, True, P, G, None, rank=rank, world=world)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_rank():
    from noisyquantumsimulator_b200 import sharding
    P, G = _population()
    ref = C.HostsimEvaluator().evaluate_host(3, True, P, G)[0]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        assert np.array_equal(got[r], ref)


def test_assignment_balances_work():
    from noisyquantumsimulator_b200 import sharding
    P, G = _population(1000)
    w = sharding.estimated_work(P, G)
    owner = sharding.shard_assignment(w, 8)
    loads = np.array([w[owner == r].sum() for r in range(8)])
    assert loads.max() / loads.min() < 1.02
    assert np.bincount(owner, minlength=8).max() - np.bincount(owner, minlength=8).min() <= 1
    # estimate tracks the true matvec count the kernel reports
    out = C.HostsimEvaluator().evaluate_host(3, True, P[:50], G[:50])[0]
    ratio = out[:, C.OUT["steps"]] / w[:50]
    assert ratio.std() / ratio.mean() < 0.05


def _sim_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import _cases as C2
    from noisyquantumsimulator_b200 import engine, sharding
    engine._default[0] = C2.HostsimEvaluator()
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kw, names = _sim_inputs()
    res = sharding.simulate_sharded("jp", 13, rank=rank, world=world, outputs=names, **kw)
    q.put((rank, {k: np.array(v) for k, v in res.columns.items()}))
    dist.barrier()
    dist.destroy_process_group()


def _sim_inputs():
    import bench
    rng = np.random.default_rng(3)
    ot = rng.uniform(5, 30, 13)
    kw = dict(bench.MEDIUM, omega_tau=ot, switching_times=np.sort(rng.uniform(0.05, 0.95, (13, 4)), axis=1) * ot[:, None],
              phases=rng.uniform(-np.pi, np.pi, (13, 5)), spacing_factor=rng.uniform(2.5, 3.2, 13))
    return kw, ("avg_fidelity", "f11", "cz_phase_fidelity", "loss_probability", "tau_total", "V", "Omega")


def test_simulate_sharded_three_ranks_equals_single_rank(monkeypatch):
    """The sharded form of simulate_CZ_gate_batch (config 4: one DE generation over the GPUs of a box): a ragged
    13-point batch over 3 gloo ranks returns, on every rank, exactly the single-rank columns."""
    from noisyquantumsimulator_b200 import engine, sharding
    monkeypatch.setitem(engine._default, 0, C.HostsimEvaluator())
    kw, names = _sim_inputs()
    ref = sharding.simulate_sharded("jp", 13, outputs=names, **kw)          # world = 1: the plain batch call
    assert sharding.block_range(13, 2, 3) == (10, 13) and sharding.block_range(13, 3, 4) == (12, 13)
    assert sharding.block_range(2, 3, 4) == (2, 2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 27500 + os.getpid() % 2000
    procs = [ctx.Process(target=_sim_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(3))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(3):
        for name in names:
            assert np.array_equal(got[r][name], ref.columns[name]), (r, name)
