"""Calibration files for the layers above the micro-physics (reference: calibration_data/README.md).

The reference describes, but never writes, ``calibration_data/neutral_atoms/rydberg_cz/n<NN>_*.json``:
"JSON files containing parameter values used, resulting error rates, gate durations, metadata (date,
simulation details)", to be "loaded by the primitives layer during initialization".  This module is
that writer/loader for batches evaluated on the GPU: one file per Rydberg level and protocol, one
record per parameter point.
"""
from __future__ import annotations

import datetime
import json
import os
from typing import Dict, List, Optional

import numpy as np

from .simulation import BatchResult

SCHEMA = "rydberg_cz_calibration/1"


def calibration_records(res: BatchResult, parameters: Dict[str, np.ndarray]) -> List[Dict]:
    """One record per point: the swept parameter values, error rates and the gate duration."""
    B = len(res.avg_fidelity)
    cols = {k: np.broadcast_to(np.asarray(v, dtype=float), (B,)) for k, v in parameters.items()}
    out = []
    for i in range(B):
        F = float(res.avg_fidelity[i])
        out.append({
            "parameters": {k: float(v[i]) for k, v in cols.items()},
            "error_rates": {"gate_infidelity": 1.0 - F, "loss_probability": float(res.loss_probability[i]),
                            "rydberg_residual": float(res.rydberg_residual[i]),
                            "phase_error_rad": float(res.phase_error_from_pi_rad[i]),
                            "infidelity_by_input": {lab: 1.0 - float(res.fidelities[i, k])
                                                    for k, lab in enumerate(("00", "01", "10", "11"))}},
            "gate_duration_s": float(res.tau_total[i]),
            "V_over_Omega": float(res.V_over_Omega[i]),
            "Omega_rad_per_s": float(res.Omega[i]),
        })
    return out


def write_cz_calibration(root: str, res: BatchResult, parameters: Dict[str, np.ndarray], *, protocol: str,
                         n_rydberg: int, species: str = "Rb87", tag: str = "sweep", hilbert_space_dim: int = 3,
                         include_noise: bool = True, n_shots: int = 0, extra: Optional[Dict] = None) -> str:
    """Write ``<root>/neutral_atoms/rydberg_cz/n<NN>_<protocol>_<tag>.json`` and return its path."""
    d = os.path.join(root, "neutral_atoms", "rydberg_cz")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, f"n{int(n_rydberg)}_{protocol}_{tag}.json")
    doc = {
        "schema": SCHEMA,
        "metadata": {"date": datetime.datetime.now(datetime.timezone.utc).isoformat(timespec="seconds"),
                     "generator": "noisyquantumsimulator_b200 (rydcz CUDA path)", "species": species,
                     "n_rydberg": int(n_rydberg), "protocol": protocol, "hilbert_space_dim": hilbert_space_dim,
                     "include_noise": bool(include_noise), "monte_carlo_shots": int(n_shots),
                     "n_points": len(res.avg_fidelity), **(extra or {})},
        "records": calibration_records(res, parameters),
    }
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
    return path


def load_cz_calibration(path: str) -> Dict:
    """Load a calibration file; arrays of the common fields are added under ``"arrays"``."""
    with open(path) as f:
        doc = json.load(f)
    if doc.get("schema") != SCHEMA:
        raise ValueError(f"{path}: not a {SCHEMA} file")
    rec = doc["records"]
    doc["arrays"] = {"gate_infidelity": np.array([r["error_rates"]["gate_infidelity"] for r in rec]),
                     "loss_probability": np.array([r["error_rates"]["loss_probability"] for r in rec]),
                     "gate_duration_s": np.array([r["gate_duration_s"] for r in rec])}
    return doc


__all__ = ["write_cz_calibration", "load_cz_calibration", "calibration_records", "SCHEMA"]
