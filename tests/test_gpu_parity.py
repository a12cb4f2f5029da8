"""GPU parity: the CUDA path through the C ABI (libspock_b200.so on a B200) against the CPU oracle on the same
seeded inputs, against the reference's golden traces, and -- at sizes the oracle cannot reach quickly -- through
size-independent properties.  fp64 tolerance from BASELINE.json north_star: 1e-9 relative after K passes."""
import numpy as np
import pytest

from oracle.oracle import CP, SP, Oracle
from spock_b200 import (LOG, SolverOptions, build_model, get_random_system_specs, get_server_heat_specs,
                        get_test_fixture_specs, solve_model_)
from spock_b200 import model as M
from spock_b200.handle import Handle
from spock_b200.problem import flatten_problem
from tests.helpers import SMALL_PROBLEMS, check_k_steps, check_operators, flat_of, golden_cp, golden_sp, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("factor_mode", [0, 1])
@pytest.mark.parametrize("name", list(SMALL_PROBLEMS))
def test_operators(cuda_lib, name, factor_mode):
    check_operators(cuda_lib, flat_of(name), factor_mode)


@pytest.mark.parametrize("spec", [
    ("heat", 9, 20, 2), ("heat", 6, 10, 3), ("heat", 5, 10, 4), ("random", 8, 2, 50, 20), ("random", 6, 3, 33, 7),
])
def test_operators_medium(cuda_lib, spec):
    if spec[0] == "heat":
        flat = flatten_problem(*get_server_heat_specs(spec[1], spec[2], spec[3]))
    else:
        flat = flatten_problem(*get_random_system_specs(spec[1], spec[2], spec[3], spec[4]))
    check_operators(cuda_lib, flat, 0, seed=5)
    check_operators(cuda_lib, flat, 1, seed=6)


@pytest.mark.parametrize("name,alg,K", [
    ("fixture_N3_d2", CP, 200), ("heat_N7_d2_nx5", CP, 500), ("heat_N4_d3_nx3", CP, 200),
    ("l2ball_N4_d3_nx3", CP, 200), ("random_N3_d4_nx35_nu17", CP, 100), ("pernode_cost_tv_N4_d2_nx3", CP, 150),
    ("pernode_cost_tv_N4_d2_nx3", SP, 15),
    ("heat_N7_d2_nx5", SP, 30), ("fixture_N3_d2", SP, 20), ("l2ball_N4_d3_nx3", SP, 15),
    ("nonuniform_avar_N5_d2_nx4", CP, 150), ("nonuniform_avar_N4_d3_nx3", CP, 100), ("nonuniform_avar_N5_d2_nx4", SP, 15),
])
@pytest.mark.parametrize("factor_mode", [0, 1])
def test_k_step_iterates(cuda_lib, name, alg, K, factor_mode):
    check_k_steps(cuda_lib, flat_of(name), alg, K, factor_mode=factor_mode)


def test_cp_golden_trace_full(cuda_lib):
    """All 3159 CP iterations of examples/server_heat/residuals.jl: same stopping pass, xi within 1e-9."""
    gold = golden_cp()
    h = Handle(flat_of("heat_N7_d2_nx5"), CP, lib=cuda_lib)
    res, log = h.solve(0.1 * np.ones(5), tol=1e-5, verbose_log=True)
    assert res["passes"] == 3159 and res["converged"] == 1 and res["printed_iter"] == 3158
    assert np.max(np.abs(log["xi"] - gold) / gold) < 1e-9
    o = Oracle(flat_of("heat_N7_d2_nx5"), CP)
    o.solve(0.1 * np.ones(5), tol=1e-5, verbose_log=True)
    z, v = h.get_iterate()
    assert relerr(z, o.z) < 1e-9 and relerr(v, o.v) < 1e-9


def test_sp_golden_trace(cuda_lib):
    xi, bc = golden_sp()
    h = Handle(flat_of("heat_N7_d2_nx5"), SP, lib=cuda_lib)
    res, log = h.solve(0.1 * np.ones(5), tol=1e-5, verbose_log=True)
    assert np.max(np.abs(log["xi"][:30] - xi[:30]) / xi[:30]) < 1e-9
    assert np.array_equal(log["backtrack_count"][:60], bc[:60])
    assert res["converged"] == 1 and log["xi"][-1] <= 1e-5  # converged solution within the solver tolerance


def test_silent_solve_matches_oracle(cuda_lib):
    flat = flat_of("heat_N7_d2_nx5")
    h, o = Handle(flat, CP, lib=cuda_lib), Oracle(flat, CP)
    r1, _ = h.solve(0.1 * np.ones(5), tol=1e-3)
    r2, _ = o.solve(0.1 * np.ones(5), tol=1e-3)
    assert r1["passes"] == r2["passes"] and abs(r1["xi1"] - r2["xi1"]) < 1e-9 * r2["xi1"]
    z, _ = h.get_iterate()
    assert relerr(z, o.z) < 1e-9


def test_reference_api_and_its_tests(cuda_lib):
    """The reference's own test-suite (test/model_algorithms/prox.jl, test/model_dynamics/implicit_L.jl) through
    the host mirror of build_model / solve_model!."""
    scen_tree, cost, dynamics, rms, constraints = get_test_fixture_specs()
    cp_model = build_model(scen_tree, cost, dynamics, rms, constraints, SolverOptions(M.CP, None))
    solve_model_(cp_model, [0.1, 0.1], quiet=True)
    z = cp_model.solver_state.z
    si = cp_model.solver_state_internal
    x, u, s, tau, y = (z[si.x_inds].reshape(7, 2), z[si.u_inds], z[si.s_inds], z[si.tau_inds],
                       z[si.y_inds].reshape(3, 5))
    A, B = dynamics.A, dynamics.B
    for j in range(2, 8):  # implicit_L.jl:59-66
        i, w = (j - 2) // 2 + 1, j % 2
        assert np.allclose(x[j - 1], A[w] @ x[i - 1] + B[w][:, 0] * u[i - 1])
    for i in range(1, 4):  # implicit_L.jl:70-78
        ch = scen_tree.child_mapping[i]
        assert np.allclose(rms[i - 1].E.T @ y[i - 1], [tau[c - 2] + s[c - 1] for c in ch])
    rng = np.random.default_rng(0)
    for _ in range(10):  # prox.jl:53-67
        z1, z2 = rng.random(cp_model.solver_state.nz), rng.random(cp_model.solver_state.nz)
        p1, p2 = z1.copy(), z2.copy()
        M.prox_f_(cp_model, p1, 0.1)
        M.prox_f_(cp_model, p2, 0.1)
        assert (p1 - p2) @ (z1 - z2) >= np.linalg.norm(p1 - p2) ** 2 * (1 - 1e-12)
    for _ in range(20):  # prox.jl:69-81
        v1, v2 = rng.random(cp_model.solver_state.nv), rng.random(cp_model.solver_state.nv)
        p1, p2 = v1.copy(), v2.copy()
        M.prox_h_conj_(cp_model, p1, 0.1)
        M.prox_h_conj_(cp_model, p2, 0.1)
        assert (p1 - p2) @ (v1 - v2) >= np.linalg.norm(p1 - p2) ** 2 * (1 - 1e-12)
    # LOG mode + SP through the same API
    sp_model = build_model(*get_server_heat_specs(7, 5, 2), SolverOptions(M.SP, None))
    solve_model_(sp_model, 0.1 * np.ones(5), tol=1e-5, verbose=LOG, quiet=True)
    assert sp_model.last_result["converged"] == 1


@pytest.mark.parametrize("factor_mode", [0, 1])
def test_large_tree_properties(cuda_lib, factor_mode):
    """Size-independent properties at a size the oracle does not visit here (N=14, nx=20: 16383 nodes):
    adjointness, idempotence of the projections, dynamics feasibility of S1, firm non-expansiveness,
    per-node == per-stage factors."""
    flat = flatten_problem(*get_random_system_specs(14, 2, 20, 8))
    h = Handle(flat, CP, factor_mode=factor_mode, lib=cuda_lib)
    rng = np.random.default_rng(11)
    z, v = rng.standard_normal(h.nz), rng.standard_normal(h.nv)
    lhs, rhs = float(h.L(z) @ v), float(z @ h.Lt(v))
    assert abs(lhs - rhs) < 1e-9 * abs(lhs)
    x0 = rng.standard_normal(flat.nx)
    h.set_x0(x0)
    p = h.prox_f(z, 0.2)
    nx, nu, n, nnl = flat.nx, flat.nu, flat.n, flat.n_non_leaf
    x, u = p[: n * nx].reshape(n, nx), p[n * nx: n * nx + nnl * nu].reshape(nnl, nu)
    A = [flat.A[k * nx * nx:(k + 1) * nx * nx].reshape(nx, nx).T for k in range(2)]
    B = [flat.B[k * nx * nu:(k + 1) * nx * nu].reshape(nu, nx).T for k in range(2)]
    j = np.arange(2, n + 1)
    anc, w = (j - 2) // 2, j % 2
    for mode in (0, 1):
        sel = w == mode
        pred = x[anc[sel]] @ A[mode].T + u[anc[sel]] @ B[mode].T
        assert np.max(np.abs(pred - x[j[sel] - 1])) < 1e-10 * max(1.0, np.max(np.abs(x)))
    assert np.allclose(x[0], x0)
    xu = p[: h.z_off[2]]
    assert relerr(h.projection_S1(xu), xu) < 1e-9
    q = h.prox_h_conj(v, 0.3)
    z2 = rng.standard_normal(h.nz)
    p2 = h.prox_f(z2, 0.2)
    assert (p - p2) @ (z - z2) >= np.linalg.norm(p - p2) ** 2 * (1 - 1e-10)
    v2 = rng.standard_normal(h.nv)
    q2 = h.prox_h_conj(v2, 0.3)
    assert (q - q2) @ (v - v2) >= np.linalg.norm(q - q2) ** 2 * (1 - 1e-10)
    # Moreau: v = prox_{sigma h*}(v) + sigma * Proj(v/sigma +- 1/2 shifts): the dual part is in the polar set, i.e.
    # prox_h* is idempotent-free but 1-Lipschitz; check non-expansiveness
    assert np.linalg.norm(q - q2) <= np.linalg.norm(v - v2) * (1 + 1e-12)
    h.close()


def test_large_tree_matches_oracle_k_steps(cuda_lib):
    """N=11, nx=nu=10 (2047 nodes): 50 CP passes against the oracle in both factor modes."""
    flat = flatten_problem(*get_server_heat_specs(11, 10, 2))
    for fm in (0, 1):
        check_k_steps(cuda_lib, flat, CP, 50, factor_mode=fm)


@pytest.mark.parametrize("name", ["fixture_N3_d2", "heat_N4_d3_nx3", "random_N3_d4_nx35_nu17", "heat_dense_N5_d2_nx4"])
@pytest.mark.parametrize("factor_mode", [0, 1])
def test_fused_sweeps_equal_modular_kernels(cuda_lib, name, factor_mode, monkeypatch):
    """One operator evaluation by the fused family kernels (TMA-staged factors) == the modular kernels, and the
    residuals that come out of the forward sweep == the ones from the explicit L / L' passes."""
    flat = flat_of(name)
    rng = np.random.default_rng(3)
    monkeypatch.setenv("SPK_DISABLE_FUSED", "1")
    hm = Handle(flat, CP, factor_mode=factor_mode, lib=cuda_lib)
    monkeypatch.setenv("SPK_DISABLE_FUSED", "0")
    hf = Handle(flat, CP, factor_mode=factor_mode, lib=cuda_lib)
    z0, v0, x0 = rng.standard_normal(hm.nz), rng.standard_normal(hm.nv), rng.standard_normal(flat.nx)
    for h in (hm, hf):
        h.set_x0(x0)
        h.set_iterate(z0, v0)
        h.step(3)
    (zm, vm), (zf, vf) = hm.get_iterate(), hf.get_iterate()
    assert relerr(zf, zm) < 1e-12 and relerr(vf, vm) < 1e-12
    xm, xf = hm.residuals(), hf.residuals()
    assert abs(xm[0] - xf[0]) < 1e-11 * xm[0] and abs(xm[1] - xf[1]) < 1e-11 * xm[1]
    assert hf.launch_count() < hm.launch_count()


def test_subtree_partition_over_gpus():
    """1-GPU vs G-GPU equality of the CP iterate (NCCL exchange of the cut-level entries); needs >= 2 GPUs."""
    import os
    import subprocess
    import sys

    import torch

    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if g < 4 else 4
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
         "--master-port", "29517", os.path.join(root, "tools", "multigpu_check.py")],
        capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0 and "MULTIGPU_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]


@pytest.mark.parametrize("nx,nu,d,N", [(50, 20, 2, 6), (50, 50, 2, 5), (20, 20, 2, 7), (10, 10, 4, 5), (10, 10, 2, 8),
                                       (5, 5, 2, 9), (20, 20, 3, 5), (10, 10, 3, 6)])
@pytest.mark.parametrize("factor_mode", [0, 1])
def test_shape_specialised_kernels_match_oracle(cuda_lib, nx, nu, d, N, factor_mode, monkeypatch):
    """Every compile-time-specialised instantiation of the S1 / dual kernels (spk_solver.cu g_shapes) against the
    oracle after K passes, and against the runtime-size instantiation of the same kernels."""
    flat = flatten_problem(*get_random_system_specs(N, d, nx, nu))
    K = 25
    o = Oracle(flat, CP)
    o.set_x0(0.1 * np.ones(nx))
    o.step(K)
    res = []
    for generic in ("0", "1"):
        monkeypatch.setenv("SPK_GENERIC_KERNELS", generic)
        h = Handle(flat, CP, factor_mode=factor_mode, lib=cuda_lib)
        h.set_x0(0.1 * np.ones(nx))
        h.step(K)
        res.append(h.get_iterate())
        h.close()
    (zs, vs), (zg, vg) = res
    assert relerr(zs, o.z) < 1e-9 and relerr(vs, o.v) < 1e-9
    assert relerr(zs, zg) < 1e-12 and relerr(vs, vg) < 1e-12


def test_sp_uses_fused_operator_and_matches_oracle_medium(cuda_lib):
    """SuperMann on a medium tree (N=9, nx=nu=10): 12 passes against the oracle (SP parity is defined for K <= 30)."""
    flat = flatten_problem(*get_server_heat_specs(9, 10, 2))
    check_k_steps(cuda_lib, flat, SP, 12, factor_mode=1)
    check_k_steps(cuda_lib, flat, SP, 12, factor_mode=0)


def test_cuda_graph_replay_equals_stream_launches(cuda_lib, monkeypatch):
    """The cached CUDA graphs of the operator evaluation give bitwise the same iterate as plain stream launches (CP and SP)."""
    flat = flat_of("heat_N7_d2_nx5")
    for alg, K in ((CP, 60), (SP, 15)):
        out = []
        for ng in ("0", "1"):
            monkeypatch.setenv("SPK_NO_GRAPH", ng)
            h = Handle(flat, alg, factor_mode=1, lib=cuda_lib)
            h.set_x0(0.1 * np.ones(5))
            h.step(K)
            out.append(h.get_iterate() + (h.residuals(),))
            h.close()
        assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and out[0][2] == out[1][2]


def _per_node_diagonal_cost_flat(N, d, nx, nu, uniform_terminal=False):
    """Random system with a different Diagonal cost at every node (Vector{Diagonal} in the reference, cost.jl:3-7)."""
    from spock_b200.problem import Cost, Diagonal

    scen_tree, _, dynamics, rms, constraints = get_random_system_specs(N, d, nx, nu)
    rng = np.random.default_rng(21)
    Q = [Diagonal(0.05 + rng.random(nx)) for _ in range(scen_tree.n - 1)]
    R = [Diagonal(0.5 + rng.random(nu)) for _ in range(scen_tree.n - 1)]
    QN = Diagonal(0.1 * np.ones(nx)) if uniform_terminal else [Diagonal(0.05 + rng.random(nx)) for _ in range(scen_tree.n_leaf_nodes)]
    return flatten_problem(scen_tree, Cost(Q, R, QN), dynamics, rms, constraints)


@pytest.mark.parametrize("nx,nu,d,N", [(50, 20, 2, 5), (10, 10, 4, 4), (5, 5, 2, 7), (20, 20, 3, 4)])
@pytest.mark.parametrize("uniform_terminal", [False, True])
def test_register_dual_kernel_per_node_diagonal_costs(cuda_lib, nx, nu, d, N, uniform_terminal):
    """spk_k_dual3 with per-node Diagonal sqrt-costs (UC = false instantiation) and the mixed per-node / uniform case
    (falls back to the batched kernel) against the oracle, both factor modes (per-stage: tensor-pipe S1 sweeps)."""
    flat = _per_node_diagonal_cost_flat(N, d, nx, nu, uniform_terminal)
    for fm in (0, 1):
        check_k_steps(cuda_lib, flat, CP, 20, factor_mode=fm)
    check_k_steps(cuda_lib, flat, SP, 8, factor_mode=1)


@pytest.mark.parametrize("nx,nu,d,N", [(50, 20, 2, 8), (50, 50, 2, 6), (20, 20, 2, 9), (10, 10, 4, 6), (10, 10, 2, 10),
                                       (5, 5, 2, 11), (20, 20, 3, 6), (10, 10, 3, 7)])
def test_tensor_pipe_sweeps_equal_family_kernels(cuda_lib, nx, nu, d, N, monkeypatch):
    """Per-stage factors: the DMMA level kernels (spk_mm.cuh; tiles of 8 families, ragged last tiles on every level of a
    d = 3 tree) against the CTA-per-family kernels on random iterates, and both against the oracle's projection_S1."""
    flat = flatten_problem(*get_random_system_specs(N, d, nx, nu))
    rng = np.random.default_rng(17)
    out = []
    monkeypatch.setenv("SPK_MM_MIN_COUNT", "1")  # by default levels with < 512 families use the CTA-per-family kernels
    for no_mm, no_top in (("1", "0"), ("0", "0"), ("0", "1")):
        monkeypatch.setenv("SPK_NO_MM", no_mm)
        monkeypatch.setenv("SPK_NO_TOP", no_top)
        h = Handle(flat, CP, factor_mode=1, lib=cuda_lib)
        if not out:
            z0, v0, x0 = rng.standard_normal(h.nz), rng.standard_normal(h.nv), rng.standard_normal(nx)
        h.set_x0(x0)
        h.set_iterate(z0, v0)
        h.step(4)
        out.append(h.get_iterate() + (h.residuals(),))
        h.close()
    (zf, vf, xf), (zm, vm, xm), (zl, vl, xl) = out
    # the single-CTA top-of-tree kernel does the same arithmetic per tile as the per-level launches
    assert np.array_equal(zm, zl) and np.array_equal(vm, vl) and xm == xl
    assert relerr(zm, zf) < 1e-12 and relerr(vm, vf) < 1e-12
    assert abs(xm[0] - xf[0]) < 1e-10 * xf[0] and abs(xm[1] - xf[1]) < 1e-10 * xf[1]
    o = Oracle(flat, CP)
    o.set_x0(x0)
    o.z[:], o.v[:] = z0, v0
    o.step(4)
    assert relerr(zm, o.z) < 1e-11 and relerr(vm, o.v) < 1e-11
    # default policy: tensor-pipe kernels on the wide levels, CTA-per-family kernels on the narrow ones
    monkeypatch.setenv("SPK_MM_MIN_COUNT", "16")
    h = Handle(flat, CP, factor_mode=1, lib=cuda_lib)
    h.set_x0(x0)
    h.set_iterate(z0, v0)
    h.step(4)
    zh, vh = h.get_iterate()
    h.close()
    assert relerr(zh, o.z) < 1e-11 and relerr(vh, o.v) < 1e-11


def test_state_checkpoint_resume_is_bit_exact(cuda_lib):
    """spk_get_state / spk_set_state + the iterate = everything the next pass reads: a SuperMann run interrupted after
    9 passes and resumed in a FRESH handle reproduces the uninterrupted run bit for bit (Anderson memory included)."""
    flat = flat_of("heat_N7_d2_nx5")
    for alg, k1, k2 in ((SP, 9, 11), (CP, 20, 15)):
        ref = Handle(flat, alg, factor_mode=1, lib=cuda_lib)
        ref.set_x0(0.1 * np.ones(5))
        ref.step(k1)
        z1, v1 = ref.get_iterate()
        st = ref.get_state()
        ref.step(k2)
        z_ref, v_ref = ref.get_iterate()
        ref.close()
        h = Handle(flat, alg, factor_mode=1, lib=cuda_lib)
        h.set_x0(0.1 * np.ones(5))
        assert h.get_state().shape == st.shape == ((12 * (h.nz + h.nv) + 9,) if alg == SP else (h.nz + h.nv,))
        h.set_iterate(z1, v1)
        h.set_state(st)
        h.step(k2)
        z, v = h.get_iterate()
        h.close()
        # (r_safe and the iteration counters are per-solve locals, sp.jl:366-386: both runs restart them at the split)
        assert np.array_equal(z, z_ref) and np.array_equal(v, v_ref)


def test_power_iteration_L_norm(cuda_lib):
    """||L||^2 from the device power iteration == the largest eigenvalue of L'L assembled column by column from the oracle's L;
    the reference's hard-coded values (10 / 3.28) are upper bounds of it for server_heat (cpock.jl:62, spock.jl:63)."""
    flat = flat_of("fixture_N3_d2")
    h, o = Handle(flat, CP, lib=cuda_lib), Oracle(flat, CP)
    Lmat = np.stack([o.L(e) for e in np.eye(h.nz)], axis=1)
    lam = float(np.linalg.eigvalsh(Lmat.T @ Lmat)[-1])
    est = h.estimate_L_norm(400)
    assert abs(est - lam) < 1e-6 * lam
    h.close()
    hh = Handle(flat_of("heat_N7_d2_nx5"), CP, lib=cuda_lib)
    est = hh.estimate_L_norm(200, apply=True)
    assert 1.0 < est < 3.28
    res, _ = hh.solve(0.1 * np.ones(5), tol=1e-3)  # converges with the estimated (larger) step sizes
    assert res["converged"] == 1
    hh.close()


@pytest.mark.parametrize("alg,verbose_log", [(CP, False), (SP, False), (SP, True)])
@pytest.mark.parametrize("name", ["heat_N7_d2_nx5", "l2ball_N4_d3_nx3", "random_N5_d2_nx6_nu3"])
def test_device_resident_loop_equals_host_driven_loop(cuda_lib, alg, verbose_log, name, monkeypatch):
    """solve_model! with the loop scalars on the device (termination test, Anderson / SuperMann decisions, polled every
    8 passes) against the host-driven loop (one synchronisation per decision): same stopping pass, same per-pass log,
    same branch counts, iterate within rounding; and a warm-started second solve from the persisted state."""
    flat = flat_of(name)
    x0 = 0.1 * np.ones(flat.nx)
    out = []
    for host_loop in ("1", "0"):
        monkeypatch.setenv("SPK_HOST_LOOP", host_loop)
        h = Handle(flat, alg, factor_mode=1, lib=cuda_lib)
        r1, log1 = h.solve(x0, tol=1e-4, verbose_log=verbose_log)
        z1, v1 = h.get_iterate()
        r2, log2 = h.solve(0.9 * x0, tol=1e-4, verbose_log=verbose_log)  # warm start (cpock.jl:180-186)
        z2, v2 = h.get_iterate()
        out.append((r1, log1, z1, v1, r2, log2, z2, v2))
        h.close()
    (a1, la1, za1, va1, a2, la2, za2, va2), (b1, lb1, zb1, vb1, b2, lb2, zb2, vb2) = out
    if alg == CP:  # deterministic contraction: identical stopping pass, log and iterate
        for ra, rb in ((a1, b1), (a2, b2)):
            assert ra["passes"] == rb["passes"] and ra["converged"] == rb["converged"] == 1 and ra["printed_iter"] == rb["printed_iter"]
            assert abs(ra["xi1"] - rb["xi1"]) <= 1e-9 * abs(ra["xi1"])
        for la, lb in ((la1, lb1), (la2, lb2)):
            assert np.allclose(la["xi1"], lb["xi1"], rtol=1e-9, atol=0) and np.allclose(la["xi2"], lb["xi2"], rtol=1e-9, atol=0)
        assert relerr(zb1, za1) < 1e-11 and relerr(vb1, va1) < 1e-11 and relerr(zb2, za2) < 1e-11 and relerr(vb2, va2) < 1e-11
    else:
        # SuperMann + Anderson amplify rounding differences of the reductions chaotically: the two loops
        # are compared pass by pass over the first 25 passes; both must converge to the solver tolerance
        for ra, rb in ((a1, b1), (a2, b2)):
            assert ra["converged"] == rb["converged"] == 1
            assert ra["k1_count"] + ra["k2_count"] + ra["fallback_count"] == ra["passes"]
            assert rb["k1_count"] + rb["k2_count"] + rb["fallback_count"] == rb["passes"]
        k = min(25, len(la1["xi1"]), len(lb1["xi1"]))
        assert np.array_equal(la1["backtrack_count"][:k], lb1["backtrack_count"][:k])
        assert np.allclose(la1["xi1"][:k], lb1["xi1"][:k], rtol=1e-7, atol=0) and np.allclose(la1["xi2"][:k], lb1["xi2"][:k], rtol=1e-7, atol=0)
        if verbose_log:
            assert np.allclose(la1["xi"][:k], lb1["xi"][:k], rtol=1e-7, atol=0)
            assert la1["xi"][-1] <= 1e-4 and lb1["xi"][-1] <= 1e-4 and la2["xi"][-1] <= 1e-4 and lb2["xi"][-1] <= 1e-4


def test_single_process_device_group():
    """All GPUs of the box behind ONE handle in ONE process (spk_problem.devices) == one GPU; needs >= 2 GPUs."""
    import os
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tools", "device_group_check.py")], capture_output=True, text=True,
                         timeout=900, cwd=root)
    assert res.returncode == 0 and "DEVICE_GROUP_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
