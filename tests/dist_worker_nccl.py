"""Worker of tests/test_gpu_sharded_nccl.py: launched by torchrun, one rank per GPU, NCCL.

Runs the product's two multi-GPU paths on real devices and checks them against rank 0's own single-GPU
evaluation of the whole batch:
  1. sharding.simulate_sharded  -- a bang-bang population (config 4 shape) and a sweep with Monte-Carlo shots
                                   (config 3 shape; shots keyed by the global point index);
  2. sharding.evaluate_sharded  -- evaluator rows, work-balanced deal;
  3. optimize_cz_gate(population_shards=(rank, world)) -- a short DE run, identical optimum on every rank.
Prints one line 'NCCL_SHARDED_OK world=<n>' from rank 0 on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bench
    from noisyquantumsimulator_b200 import sharding
    from noisyquantumsimulator_b200 import rydberg_gates as rg
    from noisyquantumsimulator_b200.engine import get_evaluator
    from noisyquantumsimulator_b200.rydberg_gates.simulation import METRIC_OUTPUTS
    import importlib
    ocg = importlib.import_module("noisyquantumsimulator_b200.rydberg_gates.optimize_cz_gate")

    # 1a. population of 8192 + 5 bang-bang candidates (ragged last block)
    B = 8192 + 5
    cand = bench.make_candidates(B, 42)
    kw = dict(bench.MEDIUM, **cand)
    sharded = sharding.simulate_sharded("jp", B, rank=rank, world=world, device=local, outputs=METRIC_OUTPUTS, **kw)
    assert sharded.local_failure is None
    single = rg.simulate_CZ_gate_batch("jp", B=B, device=local, outputs=METRIC_OUTPUTS, warn=False, **kw)
    for name in METRIC_OUTPUTS:
        assert np.array_equal(sharded.columns[name], single.columns[name]), name
    # 1b. sweep with Monte-Carlo shots: 512 points x 16 shots
    rng = np.random.default_rng(1234)
    n = 515
    kw3 = dict(bench.MEDIUM, laser_2_power=10 ** rng.uniform(np.log10(0.05), np.log10(2.0), n),
               delta_over_omega=rng.uniform(0.2, 0.5, n), omega_tau=rng.uniform(3.5, 5.5, n), temperature=10e-6)
    names3 = ("avg_fidelity", "f11", "loss_probability")
    sh3 = sharding.simulate_sharded("lp", n, rank=rank, world=world, device=local, n_shots=16, seed=7, outputs=names3, **kw3)
    si3 = rg.simulate_CZ_gate_batch("lp", B=n, device=local, n_shots=16, seed=7, outputs=names3, warn=False, **kw3)
    for name in names3:
        assert np.array_equal(sh3.columns[name], si3.columns[name]), name
    # 2. evaluator rows
    plan = rg.build_plan("jp", B=1024, warn=False, **dict(bench.MEDIUM, **bench.make_candidates(1024, 5)))
    ev = get_evaluator(local)
    rows = sharding.evaluate_sharded(ev, 3, True, plan.points, plan.segs, plan.n_seg, rank=rank, world=world)
    ref_rows = ev.evaluate_host(3, True, plan.points, plan.segs, plan.n_seg)[0]
    assert np.array_equal(rows[:, :14], ref_rows[:, :14])
    # 3. sharded DE generation
    a = ocg.ApparatusConstraints(spacing_factor=5.373584)
    r = ocg.optimize_cz_gate("lp", a, maxiter=3, popsize=8, verbose=False, cache=ocg.SimulationCache(),
                             population_shards=(rank, world), device=local, polish=False)
    best = torch.tensor(list(r.best_params) + [r.best_cost], dtype=torch.float64, device="cuda")
    gathered = [torch.empty_like(best) for _ in range(world)]
    dist.all_gather(gathered, best)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    if rank == 0:
        alone = ocg.optimize_cz_gate("lp", a, maxiter=3, popsize=8, verbose=False, cache=ocg.SimulationCache(),
                                     device=local, polish=False)
        assert np.array_equal(alone.best_params, r.best_params) and alone.best_cost == r.best_cost
        print(f"NCCL_SHARDED_OK world={world}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
