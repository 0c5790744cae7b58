"""Drop-in import path of the reference (``sys.path.insert(0, "src")`` then
``from qpu_simulator.micro_physics.neutral_atoms.rydberg_gates.simulation import simulate_CZ_gate``,
as examples/research_parameter_sweeps.py:38, validate_bangbang.py:16 and diagnose_bb.py:14 do).
Only the Rydberg-CZ hot path exists; it resolves to noisyquantumsimulator_b200.rydberg_gates."""
