"""Shared helpers for the parity tests (same code drives the host-emulation and the GPU runs)."""
import os

import numpy as np

from oracle.oracle import CP, SP, Oracle
from spock_b200.examples import get_random_system_specs, get_server_heat_specs, get_test_fixture_specs
from spock_b200.handle import Handle
from spock_b200.problem import flatten_problem

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cp():
    return np.loadtxt(os.path.join(GOLDEN, "residuals_cp.txt"))


def golden_sp():
    g = np.loadtxt(os.path.join(GOLDEN, "residuals_sp.txt"))
    return g[:, 0], g[:, 1].astype(int)


def relerr(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(float(np.max(np.abs(b))), 1e-300))


def _per_node_cost_specs():
    """N=4, d=2, nx=3, nu=2 with a DIFFERENT dense SPD cost matrix at every node (the reference's Vector{Matrix} cost,
    cost.jl:3-7 / implicit_l.jl:230,243,290) and the total-variation risk measure (risk_measures.jl:138-147, F != 0)."""
    from spock_b200.problem import Cost, get_uniform_rms_tv_v2

    scen_tree, _, dynamics, _, constraints = get_random_system_specs(4, 2, 3, 2)
    rng = np.random.default_rng(7)

    def spd(k):
        M = rng.standard_normal((k, k))
        return M @ M.T / k + 0.1 * np.eye(k)

    cost = Cost([spd(3) for _ in range(scen_tree.n - 1)], [spd(2) for _ in range(scen_tree.n - 1)],
                [spd(3) for _ in range(scen_tree.n_leaf_nodes)])
    rms = get_uniform_rms_tv_v2([0.4, 0.6], 0.2, 2, 4)
    return scen_tree, cost, dynamics, rms, constraints


SMALL_PROBLEMS = {
    "pernode_cost_tv_N4_d2_nx3": _per_node_cost_specs,
    "fixture_N3_d2": lambda: get_test_fixture_specs(),
    "heat_N7_d2_nx5": lambda: get_server_heat_specs(7, 5, 2),
    "heat_N4_d3_nx3": lambda: get_server_heat_specs(4, 3, 3),
    "heat_dense_N5_d2_nx4": lambda: get_server_heat_specs(5, 4, 2, 0.9),
    "random_N5_d2_nx6_nu3": lambda: get_random_system_specs(5, 2, 6, 3),
    "l2ball_N4_d3_nx3": lambda: get_server_heat_specs(4, 3, 3, risk="l2ball"),
    "random_N3_d4_nx35_nu17": lambda: get_random_system_specs(3, 4, 35, 17, dense_cost=True),
}


def flat_of(name):
    return flatten_problem(*SMALL_PROBLEMS[name]())


def check_operators(lib, flat, factor_mode, seed=0, tol=1e-12):
    """Every operator of the inner interface against the oracle on random vectors."""
    rng = np.random.default_rng(seed)
    o = Oracle(flat, CP)
    h = Handle(flat, CP, factor_mode=factor_mode, lib=lib)
    assert (h.nz, h.nv, h.n) == (o.nz, o.nv, o.n)
    zo, vo = o.offsets()
    assert list(zo) == h.z_off and list(vo) == h.v_off
    x0 = rng.standard_normal(flat.nx)
    o.set_x0(x0)
    h.set_x0(x0)
    for scale in (1.0, 10.0):
        z = scale * rng.standard_normal(h.nz)
        v = scale * rng.standard_normal(h.nv)
        assert relerr(h.L(z), o.L(z)) <= tol
        assert relerr(h.Lt(v), o.Lt(v)) <= tol
        xu = z[: h.z_off[2]]
        assert relerr(h.projection_S1(xu), o.projection_S1(xu)) <= tol
        z2 = z[h.z_off[2] + 1:]
        assert relerr(h.projection_S2(z2), o.projection_S2(z2)) <= tol
        for g in (0.1, 0.313):
            assert relerr(h.prox_f(z, g), o.prox_f(z, g)) <= tol
            assert relerr(h.prox_h_conj(v, g), o.prox_h_conj(v, g)) <= tol
        a, b = h.dot(z, v, 2 * z, -v, 0.3, 0.2), o.dot(z, v, 2 * z, -v, 0.3, 0.2)
        assert abs(a - b) <= 1e-11 * max(1.0, abs(b))
    # adjointness <Lz, v> = <z, L'v>
    z, v = rng.standard_normal(h.nz), rng.standard_normal(h.nv)
    lhs, rhs = float(h.L(z) @ v), float(z @ h.Lt(v))
    assert abs(lhs - rhs) <= 1e-10 * max(1.0, abs(lhs))
    h.close()


def check_k_steps(lib, flat, alg, K, factor_mode=0, tol=1e-9, x0=None):
    """Primal/dual iterates after K fixed passes vs the oracle (north_star: 1e-9 relative in fp64)."""
    x0 = 0.1 * np.ones(flat.nx) if x0 is None else x0
    o = Oracle(flat, alg)
    h = Handle(flat, alg, factor_mode=factor_mode, lib=lib)
    o.set_x0(x0)
    h.set_x0(x0)
    o.step(K)
    h.step(K)
    z, v = h.get_iterate()
    ez, ev = relerr(z, o.z), relerr(v, o.v)
    h.close()
    assert ez <= tol and ev <= tol, (ez, ev)
    return ez, ev
