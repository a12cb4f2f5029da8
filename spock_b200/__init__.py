"""spock_b200 -- B200-native Chambolle-Pock / SuperMann hot path of kul-optec/spock.jl.

Host mirror of the reference's construction/solve API on top of the C ABI in include/spock_b200.h
(libspock_b200.so: hand-written sm_100a CUDA kernels).  There is no CPU fallback.
"""
from .examples import (get_random_system_specs, get_server_heat_specs, get_test_fixture_specs)  # noqa: F401
from .model import (AA, CP, CPOCK, LOG, RB, SILENT, SP, SPOCK, AlgorithmOptions, QNewtonOptions, SolverOptions,  # noqa: F401
                    VERBOSE_LEVEL, L_, L_transpose_, build_model, prox_f_, prox_h_conj_, solve_model, solve_model_,
                    spock_dot, spock_mul_, spock_norm)
from .problem import (ConvexCone, Cost, Diagonal, Dynamics, Nonnegatives, Nonpositives, Reals, RiskMeasure,  # noqa: F401
                      SecondOrderCone, UniformRectangle, Zeros, flatten_problem, get_nonuniform_rms_avar_v2, get_uniform_rms_avar_v2,
                      get_uniform_rms_l2ball_v2, get_uniform_rms_tv_v2, get_uniform_rms_v2)
from .scenario_tree import ScenarioTree, generate_scenario_tree_uniform_branching_factor_v2  # noqa: F401

__all__ = [n for n in dir() if not n.startswith("_")]
