"""B200-native Rydberg-CZ gate evaluation (the hot path of scottjones03/NoisyQuantumSimulator).

Public surface:
  rydberg_gates            host-side mirror of the reference package interface
                           (simulate_CZ_gate, input dataclasses, protocol constants, ...)
  engine.Evaluator         batched evaluation through the C ABI (include/rydcz.h)
  build.build()            nvcc build of csrc/librydcz.so for sm_100a
"""
__version__ = "0.1.0"
