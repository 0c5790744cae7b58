"""Host-side mirror of the reference package ``qpu_simulator.micro_physics.neutral_atoms.
rydberg_gates`` (RG/__init__.py:91-321) for the CZ-gate hot path.

Same names, same argument meaning, same errors; the evolution and fidelity arithmetic runs on the
B200 through include/rydcz.h.  The SciPy differential-evolution drivers (``optimize_cz_gate``,
``optimization``) are mirrored with the same entry points but evaluate a whole population per
launch.  Not mirrored (out of scope): ``visualization`` (matplotlib) and
``optimize_jp_protocols`` (superseded in the reference by ``optimize_cz_gate``).
"""
from .constants import *  # noqa: F401,F403
from .atom_database import (ATOM_DB, get_atom_properties, get_quantum_defect, effective_n,  # noqa: F401
                            get_rydberg_energy, get_C6, get_rydberg_lifetime,
                            get_default_intermediate_state)
from .configurations import (LaserParameters, TweezerParameters, EnvironmentParameters,  # noqa: F401
                             AtomicConfiguration, get_standard_rb87_config, get_standard_cs133_config,
                             TwoPhotonExcitationConfig, NoiseSourceConfig, LPSimulationInputs,
                             JPSimulationInputs, SmoothJPSimulationInputs)
from .laser_physics import (laser_E0, single_photon_rabi, two_photon_rabi, rydberg_blockade,  # noqa: F401
                            blockade_radius, CLEBSCH_GORDAN_D2, get_clebsch_gordan)
from .trap_physics import (tweezer_spacing, trap_depth, trap_frequencies, position_uncertainty,  # noqa: F401
                           anti_trap_potential, atom_loss_probability, compute_trap_dependent_noise,
                           calculate_zeeman_shift, calculate_stark_shift)
from .noise_models import (build_all_noise_operators, build_decay_operators, build_dephasing_operators,  # noqa: F401
                           build_loss_operators, build_scatter_operators, intermediate_state_scattering_rate,
                           leakage_rate_to_adjacent_states, NoiseRates, compute_noise_rates)
from .hamiltonians import (HilbertSpace, build_hilbert_space, HS3, HS4, build_laser_hamiltonian,  # noqa: F401
                           build_detuning_hamiltonian, build_interaction_hamiltonian,
                           build_full_hamiltonian)
from .pulse_shaping import (get_pulse_envelope, pulse_envelope_gaussian, pulse_envelope_cosine,  # noqa: F401
                            pulse_envelope_blackman, pulse_envelope_drag, compute_leakage_detuning,
                            spectral_leakage_factor, PULSE_SHAPES)
from .protocols import (ProtocolParameters, LPProtocolParameters, JPProtocolParameters,  # noqa: F401
                        get_lp_protocol, get_jp_protocol, LP_DEFAULT, JP_DEFAULT,
                        JP_SWITCHING_TIMES_DEFAULT, JP_PHASES_DEFAULT, LP_OMEGA_TAU_DEFAULT,
                        LP_DELTA_OVER_OMEGA_DEFAULT, LP_XI_DEFAULT, SMOOTH_JP_PARAMS,
                        get_protocol_params, compute_phase_shift_xi, LEVINE_PICHLER_PARAMS,
                        CZ_OPTIMAL_PARAMS)
from .simulation import (simulate_CZ_gate, simulate_CZ_gate_batch, compute_CZ_fidelity,  # noqa: F401
                         compute_state_fidelity, SimulationResult, BatchResult)
from .prologue import build_plan, GatePlan  # noqa: F401
from .optimize_cz_gate import (ApparatusConstraints, SimulationCache, OptimizationResult,  # noqa: F401
                               optimize_cz_gate, run_baseline, compute_cost, extract_metrics, warm_start_bounds)
from .optimization import (HardwareOptimizationResult, optimize_CZ_parameters, EvaluatedPoint,  # noqa: F401
                           ExplorationResult, explore_parameter_space, combine_explorations)
