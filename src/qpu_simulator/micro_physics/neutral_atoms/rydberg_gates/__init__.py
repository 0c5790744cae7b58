"""Alias package: every submodule the reference's callers import resolves to the B200 mirror."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(
    os.path.dirname(os.path.abspath(__file__)))))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

_impl = importlib.import_module("noisyquantumsimulator_b200.rydberg_gates")
for _name in ("constants", "atom_database", "configurations", "laser_physics", "trap_physics",
              "noise_models", "hamiltonians", "pulse_shaping", "protocols", "simulation", "prologue",
              "optimize_cz_gate", "optimization"):
    _mod = importlib.import_module(f"noisyquantumsimulator_b200.rydberg_gates.{_name}")
    sys.modules[f"{__name__}.{_name}"] = _mod
    globals()[_name] = _mod
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("_")})
