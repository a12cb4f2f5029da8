"""Workload generators: the reference's benchmark problem (examples/server_heat/server_heat.jl), its
unit-test fixture (test/model_algorithms/prox.jl:1-51) and the synthetic configurations of
BASELINE.json / SURVEY.md section 8(d).  All deterministic."""
from __future__ import annotations

import numpy as np

from .problem import (
    Cost,
    Diagonal,
    Dynamics,
    UniformRectangle,
    get_uniform_rms_avar_v2,
    get_uniform_rms_l2ball_v2,
)
from .scenario_tree import generate_scenario_tree_uniform_branching_factor_v2


def fixed_probvec(d: int) -> np.ndarray:
    """The reference draws `rand_probvec2(d)` unseeded for d > 2 (server_heat.jl:90-94); fixed here as
    SURVEY.md section 8(d) prescribes."""
    v = np.random.default_rng(1).random(d)
    return v / v.sum()


def _server_heat_dynamics(nx: int, d: int) -> Dynamics:
    """server_heat.jl:74-88."""
    A = []
    for i in range(1, d + 1):
        Ai = np.diag([1.0 + (1.0 + (j - 1) / nx) * (i - 1) / d for j in range(1, nx + 1)])
        for r in range(nx):
            if r < nx - 1:
                Ai[r, r + 1] = 0.01
            if r > 0:
                Ai[r, r - 1] = 0.01
        A.append(Ai)
    B = [np.eye(nx) for _ in range(d)]
    return Dynamics(A, B)


def _box(scen_tree, nx, nu, x_min=-1.0, x_max=1.0, u_min=-1.5, u_max=1.5) -> UniformRectangle:
    return UniformRectangle(
        x_min, x_max, u_min, u_max,
        scen_tree.n_leaf_nodes * nx, scen_tree.n_non_leaf_nodes * (nx + nu),
        nx, nu, scen_tree.n_leaf_nodes, scen_tree.n_non_leaf_nodes,
    )


def get_server_heat_specs(N: int, nx: int, d: int, alpha: float | None = None, risk: str = "avar"):
    """examples/server_heat/server_heat.jl:1-111 (3 arguments: Diagonal costs) and :113-220
    (4 arguments `alpha`: dense costs).  Returns (scen_tree, cost, dynamics, rms, constraints).
    `risk="l2ball"` swaps AV@R for the SOC ambiguity set of config 5."""
    nu = nx
    scen_tree = generate_scenario_tree_uniform_branching_factor_v2(N, d, nx, nx)
    if alpha is None:
        cost = Cost(Diagonal(np.full(nx, 1e-1)), Diagonal(np.full(nx, 1.0)), Diagonal(np.full(nx, 1e-1)))
        alpha = 0.95
    else:
        cost = Cost(np.eye(nx) * 1e-1, np.eye(nx) * 1.0, np.eye(nx) * 1e-1)
    dynamics = _server_heat_dynamics(nx, d)
    p_ref = np.array([0.3, 0.7]) if d == 2 else fixed_probvec(d)
    if risk == "avar":
        rms = get_uniform_rms_avar_v2(p_ref, alpha, d, N)
    elif risk == "l2ball":
        rms = get_uniform_rms_l2ball_v2(np.full(d, 1.0 / d), 0.1, d, N)
    else:
        raise ValueError(risk)
    return scen_tree, cost, dynamics, rms, _box(scen_tree, nx, nu)


def get_test_fixture_specs():
    """test/model_algorithms/prox.jl:1-49 == test/model_dynamics/implicit_L.jl:1-49: N=3, d=2, nx=2, nu=1
    discretised car model, AV@R(0.95), box +-1."""
    N, d, nx, nu = 3, 2, 2, 1
    scen_tree = generate_scenario_tree_uniform_branching_factor_v2(N, d, nx, nu)
    cost = Cost(np.array([[2.2, 0.0], [0.0, 3.7]]), np.array([[3.2]]), np.array([[2.2, 0.0], [0.0, 3.7]]))
    T_s = 0.1
    # Julia: [[1.,0.] [T_s, 1.0 - (i - 1) / d * T_s]] is hcat of two columns
    A = [np.array([[1.0, T_s], [0.0, 1.0 - (i - 1) / d * T_s]]) for i in range(1, d + 1)]
    B = [np.array([[0.0], [T_s]]) for _ in range(d)]
    rms = get_uniform_rms_avar_v2([0.3, 0.7], 0.95, d, N)
    return scen_tree, cost, Dynamics(A, B), rms, _box(scen_tree, nx, nu, -1.0, 1.0, -1.0, 1.0)


def get_random_system_specs(N: int, d: int, nx: int, nu: int, dense_cost: bool = False):
    """A_w = 0.9 G / rho(G), G ~ N(0,1)^{nx x nx}, B_w ~ N(0,1)^{nx x nu}/sqrt(nx),
    numpy.default_rng(2024 + w); Q = QN = 0.1 I, R = I; AV@R(p, 0.95); box x in [-1,1], u in [-1.5,1.5]."""
    scen_tree = generate_scenario_tree_uniform_branching_factor_v2(N, d, nx, nu)
    A, B = [], []
    for w in range(1, d + 1):
        rng = np.random.default_rng(2024 + w)
        G = rng.standard_normal((nx, nx))
        rho = np.max(np.abs(np.linalg.eigvals(G)))
        A.append(0.9 * G / rho)
        B.append(rng.standard_normal((nx, nu)) / np.sqrt(nx))
    if dense_cost:
        cost = Cost(np.eye(nx) * 1e-1, np.eye(nu) * 1.0, np.eye(nx) * 1e-1)
    else:
        cost = Cost(Diagonal(np.full(nx, 1e-1)), Diagonal(np.full(nu, 1.0)), Diagonal(np.full(nx, 1e-1)))
    p_ref = np.array([0.3, 0.7]) if d == 2 else fixed_probvec(d)
    rms = get_uniform_rms_avar_v2(p_ref, 0.95, d, N)
    return scen_tree, cost, Dynamics(A, B), rms, _box(scen_tree, nx, nu)
