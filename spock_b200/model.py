"""Host mirror of the reference's model API (src/model.jl, src/models/cpock.jl, src/models/spock.jl).

Same names and argument meaning as the Julia functions; Julia's `!` suffix is spelled `_`:
`solve_model!` -> `solve_model_`, `L!` -> `L_`, `prox_f!` -> `prox_f_`, ...  Vectors keep the reference
layout and live on the host as numpy arrays (`model.solver_state.z`, `.v`), synchronised with the device
around each solve exactly like the Julia shim (julia/SpockB200.jl) does.  Index vectors
(`solver_state_internal.x_inds` ...) are 0-based numpy index arrays here.

Everything numerical runs in libspock_b200.so; this module only moves data.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib
from .handle import Handle
from .problem import FlatProblem, flatten_problem
from .scenario_tree import ScenarioTree


class AlgorithmOptions(enum.IntEnum):  # model.jl:61-64
    CP = 1
    SP = 2


class QNewtonOptions(enum.IntEnum):  # model.jl:66-69
    AA = 1
    RB = 2


class VERBOSE_LEVEL(enum.IntEnum):  # model.jl:252-255
    SILENT = 1
    LOG = 2


CP, SP = AlgorithmOptions.CP, AlgorithmOptions.SP
AA, RB = QNewtonOptions.AA, QNewtonOptions.RB
SILENT, LOG = VERBOSE_LEVEL.SILENT, VERBOSE_LEVEL.LOG


@dataclass
class SolverOptions:  # model.jl:71-74
    algorithm: AlgorithmOptions = CP
    qnewton: Optional[QNewtonOptions] = None
    # extensions (not in the reference): where the Riccati factors live, and an override of the hard-coded L_norm
    factor_mode: int = _lib.FACTORS_PER_NODE
    L_norm: float = 0.0
    # estimate ||L||^2 by power iteration at build time instead of the reference's hard-coded 10.0 / 3.28
    # (cpock.jl:62-63, spock.jl:63-64: valid for its server_heat data only); ignored when L_norm > 0
    estimate_L_norm: bool = False


@dataclass
class GenericSolverState:  # model.jl:23-44 (only the fields a caller reads)
    nz: int
    nv: int
    L_norm: float
    z: np.ndarray
    v: np.ndarray
    res_0: np.ndarray = field(default_factory=lambda: np.array([-np.inf, -np.inf]))


@dataclass
class SolverStateInternal:  # model.jl:80-125: index vectors and offsets (0-based)
    x_inds: np.ndarray
    u_inds: np.ndarray
    s_inds: np.ndarray
    tau_inds: np.ndarray
    y_inds: np.ndarray
    v_offsets: list


@dataclass
class ProblemDefinition:  # model.jl:46-55
    x0: np.ndarray
    nx: int
    nu: int
    scen_tree: ScenarioTree
    rms: object
    cost: object
    dynamics: object
    constraints: object


class _Model:
    algorithm = CP

    def __init__(self, scen_tree, cost, dynamics, rms, constraints, options: SolverOptions, lib=None):
        flat: FlatProblem = flatten_problem(scen_tree, cost, dynamics, rms, constraints)
        self.flat = flat
        self.handle = Handle(flat, int(self.algorithm), L_norm=options.L_norm, factor_mode=options.factor_mode, lib=lib)
        h = self.handle
        if options.estimate_L_norm and not options.L_norm > 0:
            h.estimate_L_norm(apply=True)
        self.solver_state = GenericSolverState(h.nz, h.nv, h.L_norm, np.zeros(h.nz), np.zeros(h.nv))
        zo = h.z_off + [h.nz]
        self.solver_state_internal = SolverStateInternal(
            *(np.arange(zo[k], zo[k + 1]) for k in range(5)), v_offsets=list(h.v_off)
        )
        self.problem_definition = ProblemDefinition(np.zeros(flat.nx), flat.nx, flat.nu, scen_tree, rms, cost, dynamics,
                                                    constraints)
        self.last_result = None
        self.last_log = None


class CPOCK(_Model):  # model.jl:127-131
    algorithm = CP


class SPOCK(_Model):  # model.jl:228-232
    algorithm = SP


def build_model(scen_tree, cost, dynamics, rms, constraints, solver_options: SolverOptions = SolverOptions(CP, None),
                lib=None):
    """model.jl:273-297."""
    if solver_options.algorithm == CP:
        return CPOCK(scen_tree, cost, dynamics, rms, constraints, solver_options, lib)
    if solver_options.algorithm == SP:
        return SPOCK(scen_tree, cost, dynamics, rms, constraints, solver_options, lib)
    raise ValueError("This combination of solver options is not supported.")


def solve_model_(model: _Model, x0, tol: float = 1e-3, verbose: VERBOSE_LEVEL = SILENT, z0=None, v0=None,
                 gamma: Optional[float] = None, sigma: Optional[float] = None, max_iter: int = 0, quiet: bool = False):
    """solve_model! (cpock.jl:169-189 / spock.jl:231-251).  Returns None like the reference; the result is
    `model.solver_state.z` / `.v`.  z, v persist on the device between solves (warm start)."""
    if (z0 is None) != (v0 is None):
        # the reference's guard is a logical-or followed by copyto!(., nothing): a MethodError
        raise ValueError("z0 and v0 must be given together")
    res, log = model.handle.solve(x0, tol=tol, max_iter=max_iter, gamma=gamma or 0.0, sigma=sigma or 0.0,
                                  verbose_log=(verbose == LOG), z0=z0, v0=v0)
    model.problem_definition.x0[:] = np.asarray(x0, dtype=np.float64)
    model.solver_state.z, model.solver_state.v = model.handle.get_iterate()
    model.last_result, model.last_log = res, log
    if not quiet:
        print(f"Finished in {res['printed_iter']} iterations.")  # cp.jl:230 / sp.jl:468
    return None


solve_model = solve_model_


# ---- inner operator interface (src/model_dynamics/implicit_l.jl), in place on caller vectors like the reference
def L_(model, z, v):
    v[:] = model.handle.L(z)


def L_transpose_(model, z, v):
    z[:] = model.handle.Lt(v)


def spock_mul_(model, C, trans: bool, B, alpha: float, beta: float):
    model.handle.mul(C, trans, B, alpha, beta)


def spock_dot(model, a_z, a_v, b_z, b_v, gamma, sigma):
    return model.handle.dot(a_z, a_v, b_z, b_v, gamma, sigma)


def spock_norm(model, a_z, a_v, gamma, sigma):
    return float(np.sqrt(model.handle.dot(a_z, a_v, a_z, a_v, gamma, sigma)))


def prox_f_(model, arg, gamma):
    model.handle.set_x0(model.problem_definition.x0)
    arg[:] = model.handle.prox_f(arg, gamma)


def prox_h_conj_(model, arg, sigma):
    arg[:] = model.handle.prox_h_conj(arg, sigma)
