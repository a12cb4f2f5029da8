"""CPU-only parity of the PRODUCT's kernels and solver orchestration against the oracle, using the host-emulation
build (tests/hostemu: same sources compiled with -DSPK_HOST_EMU, one lane per group, serial).  This checks index
arithmetic and the CP/SuperMann state machines without a GPU; the GPU parity proper is tests/test_gpu_parity.py."""
import numpy as np
import pytest

from oracle.oracle import CP, SP, Oracle
from spock_b200.examples import get_random_system_specs
from spock_b200.handle import Handle
from spock_b200.problem import flatten_problem
from tests.helpers import (SMALL_PROBLEMS, check_k_steps, check_nan_never_converges, check_operators, check_soc_branches, check_sp_options,
                           check_soc_edge_vectors, check_warm_started_second_solve, flat_of, golden_cp, golden_sp, relerr)


@pytest.mark.parametrize("factor_mode", [0, 1])
@pytest.mark.parametrize("name", list(SMALL_PROBLEMS))
def test_operators(emu_lib, name, factor_mode):
    check_operators(emu_lib, flat_of(name), factor_mode)


@pytest.mark.parametrize("name", ["fixture_N3_d2", "heat_N4_d3_nx3", "l2ball_N4_d3_nx3", "random_N5_d2_nx6_nu3"])
def test_offline_factors(emu_lib, name):
    flat = flat_of(name)
    o, h = Oracle(flat, CP), Handle(flat, CP, lib=emu_lib)
    from oracle.oracle import kernel_projector
    for i in (1, 2, flat.n_non_leaf, flat.n_non_leaf + 1, flat.n):
        P, ABK, K, Rinv, ls = h.factors(i)
        Po, ABKo, Ko, Rc = o.factors(i)
        assert relerr(P, Po) < 1e-12
        if i > 1:
            assert relerr(ABK, ABKo) < 1e-12
        if i <= flat.n_non_leaf:
            assert relerr(K, Ko) < 1e-12
            assert relerr(Rinv, np.linalg.inv(Rc @ Rc.T)) < 1e-11
    assert relerr(ls, kernel_projector(flat.E2d, flat.F2d)) < 1e-12


@pytest.mark.parametrize("name,alg,K", [
    ("fixture_N3_d2", CP, 200), ("heat_N7_d2_nx5", CP, 300), ("heat_N4_d3_nx3", CP, 200),
    ("l2ball_N4_d3_nx3", CP, 200), ("pernode_cost_tv_N4_d2_nx3", CP, 150), ("pernode_cost_tv_N4_d2_nx3", SP, 15),
    ("heat_N7_d2_nx5", SP, 30), ("fixture_N3_d2", SP, 20), ("l2ball_N4_d3_nx3", SP, 15),
    ("nonuniform_avar_N5_d2_nx4", CP, 150), ("nonuniform_avar_N4_d3_nx3", CP, 100), ("nonuniform_avar_N5_d2_nx4", SP, 15),
])
@pytest.mark.parametrize("host_loop", ["0", "1"])
def test_k_step_iterates(emu_lib, name, alg, K, host_loop, monkeypatch):
    """both loop drivers: device-resident control block (default) and the host-driven loop (SPK_HOST_LOOP=1)"""
    monkeypatch.setenv("SPK_HOST_LOOP", host_loop)
    check_k_steps(emu_lib, flat_of(name), alg, K, factor_mode=K % 2)


def test_cp_golden_trace_prefix(emu_lib):
    gold = golden_cp()
    h = Handle(flat_of("heat_N7_d2_nx5"), CP, lib=emu_lib)
    res, log = h.solve(0.1 * np.ones(5), tol=1e-5, verbose_log=True, max_iter=400)
    assert res["passes"] == 400 and res["converged"] == 0
    assert np.max(np.abs(log["xi"] - gold[:400]) / gold[:400]) < 1e-9


@pytest.mark.parametrize("host_loop", ["0", "1"])
def test_sp_golden_trace_prefix(emu_lib, host_loop, monkeypatch):
    monkeypatch.setenv("SPK_HOST_LOOP", host_loop)
    xi, bc = golden_sp()
    h = Handle(flat_of("heat_N7_d2_nx5"), SP, lib=emu_lib)
    res, log = h.solve(0.1 * np.ones(5), tol=1e-5, verbose_log=True, max_iter=60)
    assert np.max(np.abs(log["xi"][:30] - xi[:30]) / xi[:30]) < 1e-9
    assert np.array_equal(log["backtrack_count"][:60], bc[:60])


@pytest.mark.parametrize("host_loop", ["0", "1"])
def test_silent_termination_and_warm_start(emu_lib, host_loop, monkeypatch):
    monkeypatch.setenv("SPK_HOST_LOOP", host_loop)
    flat = flat_of("heat_N7_d2_nx5")
    for alg in (CP, SP):
        h, o = Handle(flat, alg, lib=emu_lib), Oracle(flat, alg)
        x0 = 0.1 * np.ones(5)
        r1, _ = h.solve(x0, tol=1e-3)
        r2, _ = o.solve(x0, tol=1e-3)
        if alg == CP:
            assert r1["passes"] == r2["passes"] and r1["printed_iter"] == r2["printed_iter"]
            assert abs(r1["xi1"] - r2["xi1"]) < 1e-9 * r2["xi1"]
        assert r1["converged"] == 1
        # second solve from the stored iterate (implicit warm start, cpock.jl:169-189) needs fewer passes
        r3, _ = h.solve(1.02 * x0, tol=1e-3)
        assert r3["converged"] == 1 and r3["passes"] < r1["passes"]
        # explicit z0/v0
        r4, _ = h.solve(x0, tol=1e-3, z0=np.zeros(h.nz), v0=np.zeros(h.nv))
        if alg == CP:  # SP keeps its Anderson memory across solves like the reference
            assert r4["passes"] == r1["passes"]
        with pytest.raises(Exception):
            h.solve(x0, z0=np.zeros(h.nz))
        with pytest.raises(ValueError):
            h.L(np.zeros(h.nz + 1))  # implicit_l.jl:184-186


def test_state_checkpoint_resume(emu_lib):
    """spk_get_state / spk_set_state through the host-emulation build (host-driven loops): an interrupted SuperMann / CP run
    resumed in a fresh handle reproduces the uninterrupted one bit for bit."""
    import numpy as np

    flat = flat_of("fixture_N3_d2")
    for alg, k1, k2 in ((SP, 7, 6), (CP, 10, 8)):
        ref = Handle(flat, alg, factor_mode=1, lib=emu_lib)
        ref.set_x0(0.1 * np.ones(flat.nx))
        ref.step(k1)
        z1, v1 = ref.get_iterate()
        st = ref.get_state()
        ref.step(k2)
        z_ref, v_ref = ref.get_iterate()
        ref.close()
        h = Handle(flat, alg, factor_mode=1, lib=emu_lib)
        h.set_x0(0.1 * np.ones(flat.nx))
        assert st.shape == ((12 * (h.nz + h.nv) + 9,) if alg == SP else (h.nz + h.nv,))
        h.set_iterate(z1, v1)
        h.set_state(st)
        h.step(k2)
        z, v = h.get_iterate()
        h.close()
        assert np.array_equal(z, z_ref) and np.array_equal(v, v_ref)


def test_power_iteration_L_norm(emu_lib):
    """spk_estimate_L_norm == largest eigenvalue of L'L assembled from the oracle's L, column by column."""
    import numpy as np

    flat = flat_of("fixture_N3_d2")
    h, o = Handle(flat, CP, lib=emu_lib), Oracle(flat, CP)
    Lmat = np.stack([o.L(e) for e in np.eye(h.nz)], axis=1)
    lam = float(np.linalg.eigvalsh(Lmat.T @ Lmat)[-1])
    assert abs(h.estimate_L_norm(400) - lam) < 1e-6 * lam
    h.close()


def test_operator_calls_do_not_disturb_the_supermann_memory(emu_lib):
    """Like the reference's spock_mul! / spock_dot, spk_apply / spk_dot between two solves leave the persistent SuperMann
    state (delta_z, delta_v, Anderson columns) alone: the warm-started second solve is bit-identical with and without them."""
    import numpy as np

    flat = flat_of("heat_N4_d3_nx3")
    x0 = 0.1 * np.ones(flat.nx)
    rng = np.random.default_rng(4)
    out = []
    for poke in (False, True):
        h = Handle(flat, SP, lib=emu_lib)
        h.solve(x0, tol=1e-3)
        if poke:
            z, v = rng.standard_normal(h.nz), rng.standard_normal(h.nv)
            h.L(z), h.Lt(v), h.prox_f(z, 0.2), h.prox_h_conj(v, 0.3), h.projection_S1(z[: h.z_off[2]])
            h.dot(z, v, z, v, 0.3, 0.2)
        r, log = h.solve(0.8 * x0, tol=1e-4)
        out.append((r["passes"], log["xi1"], h.get_iterate()))
        h.close()
    assert out[0][0] == out[1][0] and np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][2][0], out[1][2][0]) and np.array_equal(out[0][2][1], out[1][2][1])


def test_build_model_with_estimated_L_norm(emu_lib):
    """SolverOptions(estimate_L_norm=True): the step sizes come from the power-iteration estimate of ||L||^2 instead of the
    reference's hard-coded constant, and the solve still converges (with fewer passes: the constant over-estimates it)."""
    import numpy as np

    from spock_b200 import SolverOptions, build_model, get_server_heat_specs, solve_model_
    from spock_b200 import model as M

    spec = get_server_heat_specs(5, 4, 2)
    x0 = 0.1 * np.ones(4)
    passes = []
    for est in (False, True):
        m = build_model(*spec, SolverOptions(M.CP, None, estimate_L_norm=est), lib=emu_lib)
        assert (m.solver_state.L_norm == 10.0) != est and 0.5 < m.solver_state.L_norm <= 10.0
        solve_model_(m, x0, tol=1e-3, quiet=True)
        assert m.last_result["converged"] == 1
        passes.append(m.last_result["passes"])
        m.handle.close()
    assert passes[1] < passes[0]


@pytest.mark.parametrize("variant", ["dual2", "dual_v1", "modular"])
@pytest.mark.parametrize("name", ["l2ball_N4_d3_nx3", "heat_N7_d2_nx5"])
def test_soc_branches(emu_lib, name, variant, monkeypatch):
    """inside / polar / general branch of the child, leaf and risk-cone SOCs through one pass (see tests/helpers.py)"""
    monkeypatch.setenv("SPK_DUAL_V1", "1" if variant == "dual_v1" else "0")
    monkeypatch.setenv("SPK_DISABLE_FUSED", "1" if variant == "modular" else "0")
    check_soc_branches(emu_lib, flat_of(name))


def test_soc_edge_vectors(emu_lib):
    check_soc_edge_vectors(emu_lib)


@pytest.mark.parametrize("alg,tol", [(CP, 1e-4), (SP, 1e-1)])
@pytest.mark.parametrize("host_loop", ["0", "1"])
def test_warm_started_second_solve_matches_oracle(emu_lib, alg, tol, host_loop, monkeypatch):
    monkeypatch.setenv("SPK_HOST_LOOP", host_loop)
    check_warm_started_second_solve(emu_lib, alg, tol)


@pytest.mark.parametrize("alg", [CP, SP])
@pytest.mark.parametrize("host_loop", ["0", "1"])
def test_nan_iterate_never_reports_convergence(emu_lib, alg, host_loop, monkeypatch):
    monkeypatch.setenv("SPK_HOST_LOOP", host_loop)
    check_nan_never_converges(emu_lib, alg)


@pytest.mark.parametrize("name", ["heat_N7_d2_nx5", "l2ball_N4_d3_nx3"])
def test_supermann_options_match_oracle(emu_lib, name):
    """working line search / K0 (non-default; the reference ships with both switched off, sp.jl:439 and sp.jl:80)"""
    from tests.helpers import flat_of
    check_sp_options(emu_lib, flat_of(name))


@pytest.mark.parametrize("name,w", [("heat_N7_d2_nx5", 1), ("heat_N4_d3_nx3", 2), ("l2ball_N4_d3_nx3", 0), ("fixture_N3_d2", 1)])
def test_shift_initialisation(emu_lib, name, w):
    from tests.helpers import check_shift_iterate, flat_of
    check_shift_iterate(emu_lib, flat_of(name), w)


def test_batched_independent_problems(emu_lib):
    from tests.helpers import check_solve_batch, flat_of
    names = ["heat_N7_d2_nx5", "random_N5_d2_nx6_nu3", "heat_N4_d3_nx3", "heat_N7_d2_nx5", "l2ball_N4_d3_nx3"]
    check_solve_batch(emu_lib, [flat_of(n) for n in names], [CP, CP, CP, SP, SP])


@pytest.mark.parametrize("name", ["fixture_N3_d2", "heat_N4_d3_nx3", "l2ball_N4_d3_nx3", "random_N5_d2_nx6_nu3", "nonuniform_avar_N5_d2_nx4",
                                  "pernode_cost_tv_N4_d2_nx3", "random_N3_d4_nx35_nu17"])
def test_offline_phase_on_device_equals_host(emu_lib, name, monkeypatch):
    from tests.helpers import check_offline_device_vs_host, flat_of
    check_offline_device_vs_host(emu_lib, flat_of(name), monkeypatch)


@pytest.mark.parametrize("N,d,nx,nu", [(2, 2, 3, 2), (2, 3, 1, 1), (3, 8, 2, 2), (3, 2, 1, 1), (4, 5, 2, 3), (2, 2, 1, 1)])
def test_edge_shapes(emu_lib, N, d, nx, nu):
    """Smallest and most lopsided trees: N = 2 (a root whose children are leaves), scalar states and inputs, the widest
    supported branching -- every operator, K passes of CP and SuperMann, both factor modes, against the oracle."""
    from tests.helpers import check_k_steps, check_operators
    flat = flatten_problem(*get_random_system_specs(N, d, nx, nu))
    for fm in (0, 1):
        check_operators(emu_lib, flat, fm)
        check_k_steps(emu_lib, flat, CP, 30, factor_mode=fm)
        check_k_steps(emu_lib, flat, SP, 10, factor_mode=fm)


@pytest.mark.parametrize("N,d,nx,nu", [(5, 2, 4, 3), (4, 3, 3, 2), (3, 4, 5, 5)])
def test_caller_supplied_mode_map(emu_lib, N, d, nx, nu):
    """spk_problem.mode_of_node: node modes that are not the generator's (j % d) + 1 -- per-node Riccati recursion on the
    device (one CTA per node, level by level), the sweeps with per-node mode lookups, against the oracle."""
    from tests.helpers import check_custom_modes, custom_mode_flat
    check_custom_modes(emu_lib, custom_mode_flat(N, d, nx, nu))
