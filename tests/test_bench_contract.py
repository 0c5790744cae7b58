"""bench.py's reference arm (CPU: the oracle restatement of the reference's mesolve path) prints the contract's
JSON line; the workload generator is seeded and inside the reference optimiser's bounds."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "cz_lindblad_gate_evals_per_sec"
    assert line["unit"] == "evals/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"] == "jp_bangbang_5seg_lindblad_3level"


def test_candidates_are_seeded_and_inside_the_optimiser_bounds():
    sys.path.insert(0, ROOT)
    import bench
    a, b = bench.make_candidates(1000, 7), bench.make_candidates(1000, 7)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert a["omega_tau"].min() >= 5.0 and a["omega_tau"].max() <= 40.0
    fr = a["switching_times"] / a["omega_tau"][:, None]
    assert np.all(np.diff(fr, axis=1) >= 0) and fr.min() >= 0.01 and fr.max() <= 0.99
    assert np.abs(a["phases"]).max() <= np.pi and a["phases"].shape == (1000, bench.N_SEG)
