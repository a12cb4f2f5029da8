"""Properties the reference's own tests check (test/model_algorithms/prox.jl, test/model_dynamics/implicit_L.jl),
run on the oracle, plus unit checks of the restated primitives."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle.oracle import CP, SP, Oracle, kernel_projector
from tests.helpers import SMALL_PROBLEMS, flat_of, relerr


def test_prox_f_firmly_nonexpansive():  # prox.jl:53-67
    o = Oracle(flat_of("fixture_N3_d2"), CP)
    o.set_x0(np.array([0.1, 0.1]))
    rng = np.random.default_rng(1)
    for _ in range(10):
        z1, z2 = rng.random(o.nz), rng.random(o.nz)
        p1, p2 = o.prox_f(z1, 0.1), o.prox_f(z2, 0.1)
        assert (p1 - p2) @ (z1 - z2) >= np.linalg.norm(p1 - p2) ** 2 * (1 - 1e-12)


def test_prox_h_conj_firmly_nonexpansive():  # prox.jl:69-81
    o = Oracle(flat_of("fixture_N3_d2"), CP)
    rng = np.random.default_rng(2)
    for _ in range(20):
        v1, v2 = rng.random(o.nv), rng.random(o.nv)
        p1, p2 = o.prox_h_conj(v1, 0.1), o.prox_h_conj(v2, 0.1)
        assert (p1 - p2) @ (v1 - v2) >= np.linalg.norm(p1 - p2) ** 2 * (1 - 1e-12)


def test_solution_feasibility():  # implicit_L.jl:59-78
    flat = flat_of("fixture_N3_d2")
    o = Oracle(flat, CP)
    res, _ = o.solve(np.array([0.1, 0.1]), tol=1e-3)
    assert res["converged"] == 1
    z = o.z.copy()
    x, u, s, tau, y = z[0:14].reshape(7, 2), z[14:17], z[17:24], z[24:30], z[30:45].reshape(3, 5)
    A = [flat.A[k * 4:(k + 1) * 4].reshape(2, 2).T for k in range(2)]
    B = [flat.B[k * 2:(k + 1) * 2] for k in range(2)]
    for j in range(2, 8):  # 1-based child j, ancestor (j-2)//2+1, mode (j % 2)+1
        i, w = (j - 2) // 2 + 1, (j % 2)
        assert np.allclose(x[j - 1], A[w] @ x[i - 1] + B[w] * u[i - 1])
    E = flat.E2d
    for i in range(1, 4):
        ch = [2 * (i - 1) + 2, 2 * (i - 1) + 3]
        assert np.allclose(E.T @ y[i - 1], [tau[c - 2] + s[c - 1] for c in ch])


def test_soc_projection_cases():  # risk_measures.jl:40-71
    o = Oracle(flat_of("fixture_N3_d2"), CP)
    inside = np.array([2.0, 1.0, 1.0])
    assert np.array_equal(o.project_soc(inside), inside)
    assert np.array_equal(o.project_soc(np.array([-2.0, 1.0, 1.0])), np.zeros(3))  # ||x|| <= -t
    p = o.project_soc(np.array([0.0, 3.0, 4.0]))
    assert np.allclose(p, [2.5, 1.5, 2.0]) and abs(np.linalg.norm(p[1:]) - p[0]) < 1e-15
    # idempotent, and the result is the nearest cone point
    q = o.project_soc(np.array([0.3, -1.0, 2.0, 0.5]))
    assert relerr(o.project_soc(q), q) < 1e-15 and np.linalg.norm(q[1:]) <= q[0] * (1 + 1e-15)


@pytest.mark.parametrize("name", list(SMALL_PROBLEMS))
def test_adjointness_and_projector(name):
    flat = flat_of(name)
    o = Oracle(flat, CP)
    rng = np.random.default_rng(3)
    z, v = rng.standard_normal(o.nz), rng.standard_normal(o.nv)
    assert abs(o.L(z) @ v - z @ o.Lt(v)) < 1e-10 * max(1.0, abs(o.L(z) @ v))
    Pk = kernel_projector(flat.E2d, flat.F2d)
    assert np.allclose(Pk @ Pk, Pk, atol=1e-13) and np.allclose(Pk, Pk.T, atol=1e-13)
    # S1 output satisfies the dynamics, S1 and S2 are idempotent
    x0 = rng.standard_normal(flat.nx)
    o.set_x0(x0)
    xu = o.projection_S1(z[: o.offsets()[0][2]])
    assert relerr(o.projection_S1(xu), xu) < 1e-10
    assert np.allclose(xu[: flat.nx], x0)
    z2 = o.projection_S2(z[o.offsets()[0][2] + 1:])
    assert relerr(o.projection_S2(z2), z2) < 1e-12


def test_per_stage_offline_equals_node_by_node():
    """riccati_offline computes one node per stage and replicates; SPO_FAITHFUL_OFFLINE=1 loops over every node."""
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from tests.helpers import flat_of; from oracle.oracle import Oracle, CP\n"
        "o = Oracle(flat_of('random_N5_d2_nx6_nu3'), CP)\n"
        "np.save(sys.argv[1], np.concatenate([np.concatenate([m.ravel() for m in o.factors(i)]) for i in range(1, o.n + 1)]))\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import tempfile
    outs = []
    for flag in ("0", "1"):
        with tempfile.NamedTemporaryFile(suffix=".npy") as f:
            subprocess.run([sys.executable, "-c", code, f.name], check=True, env=dict(os.environ, SPO_FAITHFUL_OFFLINE=flag))
            outs.append(np.load(f.name))
    assert np.array_equal(outs[0], outs[1])


def test_warm_start_guard():
    o = Oracle(flat_of("fixture_N3_d2"), CP)
    with pytest.raises(ValueError):
        o.solve(np.array([0.1, 0.1]), z0=np.zeros(o.nz))  # cpock.jl:180-183
