"""Shared helpers for the parity tests (same code drives the host-emulation and the GPU runs)."""
import os

import numpy as np
import pytest

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


def _nonuniform_rms_specs(N=5, d=2, nx=4, nu=3):
    """A different AV@R (own alpha and reference probabilities) at every non-leaf node: get_nonuniform_rms_avar_v2
    (risk_measures.jl:127-136) -- per-node b (implicit_l.jl:219,427) and per-node kernel projector (cpock.jl:71-81,125-128)."""
    from spock_b200.problem import get_nonuniform_rms_avar_v2

    scen_tree, cost, dynamics, _, constraints = get_random_system_specs(N, d, nx, nu)
    return scen_tree, cost, dynamics, get_nonuniform_rms_avar_v2(d, N, np.random.default_rng(42)), constraints


SMALL_PROBLEMS = {
    "nonuniform_avar_N5_d2_nx4": _nonuniform_rms_specs,
    "nonuniform_avar_N4_d3_nx3": lambda: _nonuniform_rms_specs(4, 3, 3, 2),
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
        for trans, Bv, Cv in ((True, v, z), (False, z, v)):  # spock_mul! with alpha, beta (implicit_l.jl:451-481)
            got = np.array(Cv, copy=True)
            h.mul(got, trans, Bv, -0.7, 1.3)
            want = -0.7 * (o.Lt(Bv) if trans else o.L(Bv)) + 1.3 * Cv
            assert relerr(got, want) <= tol
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


# ------------------------------------------------------------------------------------------------ SOC branches, warm start, NaN
# (shared by tests/test_gpu_headline_parity.py -- the CUDA path -- and tests/test_hostemu_parity.py)

def soc_mode(t, x):
    """risk_measures.jl:57-70: 0 = inside (||x|| <= t), 1 = polar (||x|| <= -t), 2 = general."""
    n = np.linalg.norm(x)
    return 0 if n <= t else (1 if n <= -t else 2)


def force_soc_branches(flat, vo, v, big=60.0):
    """Overwrites the `t` entries of the three SOC families in v so that block j takes branch j % 3:
    child SOC (t = v6_j; x = v3_j, v4_j, v5_j; implicit_l.jl:854-900), leaf SOC (t = v13; x = v11, v12; :752-810) and,
    with an SOC inside the risk cone, its t entry in v1_i (:833-845)."""
    n, nnl, nl = flat.n, flat.n_non_leaf, flat.n_leaf
    pattern = np.array([big, -big, 0.0])
    j = np.arange(n - 1)
    sel = j % 3 != 2
    v[vo[5] + j[sel]] = pattern[j[sel] % 3]
    l = np.arange(nl)
    sel = l % 3 != 2
    v[vo[9] + l[sel]] = pattern[l[sel] % 3]
    off = 0
    for ct, cd in zip(flat.cone_type, flat.cone_dim):
        if ct == 4:
            i = np.arange(nnl)
            sel = i % 3 != 2
            v[vo[0] + i[sel] * flat.ny + off] = pattern[i[sel] % 3]
        off += int(cd)
    return v


def classify_soc_after_one_pass(flat, o, z0, v0, sigma):
    """Branch taken by every SOC block of prox_h_conj! in the oracle's first pass from (z0, v0), recomputed in numpy from
    w = (v + sigma L(2 zbar - z)) / sigma with the +-1/2 shifts (implicit_l.jl:911-951)."""
    n, nnl, nl, nx, nu, ny = flat.n, flat.n_non_leaf, flat.n_leaf, flat.nx, flat.nu, flat.ny
    _, vo = o.offsets()
    zbar = o.z.copy()  # lambda = 1: z after the pass is zbar
    w = (v0 + sigma * o.L(2 * zbar - z0)) / sigma
    w[vo[4]:vo[4] + n - 1] -= 0.5
    w[vo[5]:vo[5] + n - 1] += 0.5
    w[vo[8]:vo[8] + nl] -= 0.5
    w[vo[9]:vo[9] + nl] += 0.5
    child = [soc_mode(w[vo[5] + j], np.concatenate([w[vo[2] + j * nx: vo[2] + (j + 1) * nx], w[vo[3] + j * nu: vo[3] + (j + 1) * nu],
                                                    [w[vo[4] + j]]])) for j in range(n - 1)]
    leaf = [soc_mode(w[vo[9] + l], np.concatenate([w[vo[7] + l * nx: vo[7] + (l + 1) * nx], [w[vo[8] + l]]])) for l in range(nl)]
    risk, off = [], 0
    for ct, cd in zip(flat.cone_type, flat.cone_dim):
        if ct == 4:
            risk = [soc_mode(w[vo[0] + i * ny + off], w[vo[0] + i * ny + off + 1: vo[0] + i * ny + off + cd]) for i in range(nnl)]
        off += int(cd)
    return child, leaf, risk


def check_soc_branches(lib, flat):
    """Every branch of the SOC projection, for the child, leaf and (l2-ball) risk-cone SOCs, through one full pass of the
    operator evaluation against the oracle at 1e-12 (which dual kernel runs is chosen by the caller's environment)."""
    rng = np.random.default_rng(99)
    o = Oracle(flat, CP)
    _, vo = o.offsets()
    z0 = 0.2 * rng.standard_normal(o.nz)
    v0 = force_soc_branches(flat, vo, rng.standard_normal(o.nv))
    x0 = 0.1 * np.ones(flat.nx)
    o.set_x0(x0)
    o.z[:], o.v[:] = z0, v0
    o.step(1)
    child, leaf, risk = classify_soc_after_one_pass(flat, o, z0, v0, o.default_step())
    assert set(child) == {0, 1, 2} and set(leaf) == {0, 1, 2}
    if 4 in list(flat.cone_type):
        assert set(risk) == {0, 1, 2}
    for fm in (0, 1):
        h = Handle(flat, CP, factor_mode=fm, lib=lib)
        h.set_x0(x0)
        h.set_iterate(z0, v0)
        h.step(1)
        z, v = h.get_iterate()
        h.close()
        assert relerr(z, o.z) < 1e-12 and relerr(v, o.v) < 1e-12, fm
        assert np.max(np.abs(v - o.v)) < 1e-10


def check_soc_edge_vectors(lib):
    """Explicit vectors through spk_apply(SPK_OP_PROX_HCONJ): for one child SOC and one leaf SOC each of
    inside / polar / general / zero-norm (x = 0 with t > 0, t < 0 and t = 0; risk_measures.jl:57-70 never divides by a zero
    norm because ||x|| > |t| >= 0 in the general branch) and the boundary ||x|| == t."""
    flat = flatten_problem(*get_server_heat_specs(4, 3, 3, risk="l2ball"))
    o = Oracle(flat, CP)
    _, vo = o.offsets()
    n, nl, nx, nu = flat.n, flat.n_leaf, flat.nx, flat.nu
    sigma = 0.4
    h = Handle(flat, CP, lib=lib)
    rng = np.random.default_rng(5)
    cases = [("inside", 5.0, 1.0), ("polar", -5.0, 1.0), ("general", 0.3, 1.0), ("zero_norm_pos_t", 2.0, 0.0),
             ("zero_norm_neg_t", -2.0, 0.0), ("zero_norm_zero_t", 0.0, 0.0), ("boundary", 1.0, None)]
    for name, t, scale in cases:
        w = np.zeros(o.nv)  # the vector AFTER the 1/sigma scaling and the +-1/2 shifts; mapped back to prox_h_conj!'s argument below
        j, l = 4, 2
        xc = rng.standard_normal(nx + nu + 1)
        xl = rng.standard_normal(nx + 1)
        if scale is None:
            xc, xl = xc / np.linalg.norm(xc) * t, xl / np.linalg.norm(xl) * t
        else:
            xc, xl = xc * scale, xl * scale
        w[vo[2] + j * nx: vo[2] + (j + 1) * nx] = xc[:nx]
        w[vo[3] + j * nu: vo[3] + (j + 1) * nu] = xc[nx:nx + nu]
        w[vo[4] + j], w[vo[5] + j] = xc[-1], t
        w[vo[7] + l * nx: vo[7] + (l + 1) * nx] = xl[:nx]
        w[vo[8] + l], w[vo[9] + l] = xl[-1], t
        arg = w.copy()
        arg[vo[4]:vo[4] + n - 1] += 0.5
        arg[vo[5]:vo[5] + n - 1] -= 0.5
        arg[vo[8]:vo[8] + nl] += 0.5
        arg[vo[9]:vo[9] + nl] -= 0.5
        arg *= sigma
        got, want = h.prox_h_conj(arg, sigma), o.prox_h_conj(arg, sigma)
        assert np.all(np.isfinite(got)), name
        assert np.max(np.abs(got - want)) < 1e-13 * max(1.0, np.max(np.abs(want))), name
    h.close()


def check_warm_started_second_solve(lib, alg, tol):
    """solve_model! twice on the same model (cpock.jl:169-189 / spock.jl:231-251): the second solve starts from the first
    one's (z, v) -- and, for SuperMann, from its stale Anderson memory, z_old, rz_old, delta -- with a new x0.
    Same number of passes and the iterate within 1e-9 of the oracle after both solves.  (SuperMann: a loose tolerance keeps
    both solves inside the <= 30 passes for which SP parity is defined.)"""
    flat = flatten_problem(*get_server_heat_specs(7, 5, 2))
    for fm in (0, 1):
        h, o = Handle(flat, alg, factor_mode=fm, lib=lib), Oracle(flat, alg)
        total = 0
        for x0 in (0.1 * np.ones(5), np.array([0.3, -0.2, 0.1, 0.25, -0.3])):
            rh, _ = h.solve(x0, tol=tol)
            ro, _ = o.solve(x0, tol=tol)
            assert rh["passes"] == ro["passes"] and rh["converged"] == ro["converged"] == 1
            total += ro["passes"]
            if alg == SP:
                assert total <= 30
                assert (rh["k1_count"], rh["k2_count"], rh["fallback_count"]) == (ro["k1_count"], ro["k2_count"], ro["fallback_count"])
            z, v = h.get_iterate()
            assert relerr(z, o.z) < 1e-9 and relerr(v, o.v) < 1e-9
        h.close()


def check_nan_never_converges(lib, alg):
    """A NaN in x0 makes every residual NaN; the reference's `xi <= tol` is then false and it runs to MAX_ITER
    (cp.jl:102, sp.jl:323).  The residual maxima of the product path must keep the NaN (spk_max_nn)."""
    flat = flatten_problem(*get_server_heat_specs(5, 5, 2))
    x0 = 0.1 * np.ones(5)
    x0[2] = np.nan
    for fm in (0, 1):
        h = Handle(flat, alg, factor_mode=fm, lib=lib)
        res, _ = h.solve(x0, tol=1e-3, max_iter=12)
        assert res["converged"] == 0 and res["passes"] == 12
        assert np.isnan(res["xi1"]) or np.isnan(res["xi2"])
        h.close()
    o = Oracle(flat, alg)
    ro, _ = o.solve(x0, tol=1e-3, max_iter=12)
    assert ro["converged"] == 0 and ro["passes"] == 12


def check_sp_options(lib, flat, K=12):
    """SuperMann variants behind spk_set_sp_options (working line search tau <- beta tau, sp.jl:439; K0 blind update, sp.jl:80)
    against the oracle's restatement of the same variants: iterate after K passes and the branch counts of a solve."""
    x0 = 0.1 * np.ones(flat.nx)
    for ls, k0 in ((True, False), (False, True), (True, True)):
        h, o = Handle(flat, SP, factor_mode=1, lib=lib), Oracle(flat, SP)
        h.set_sp_options(linesearch=ls, k0=k0)
        o.set_sp_options(linesearch=ls, k0=k0)
        h.set_x0(x0)
        o.set_x0(x0)
        h.step(K)
        o.step(K)
        z, v = h.get_iterate()
        assert relerr(z, o.z) < 1e-9 and relerr(v, o.v) < 1e-9, (ls, k0, relerr(z, o.z), relerr(v, o.v))
        h.close()
        h, o = Handle(flat, SP, factor_mode=0, lib=lib), Oracle(flat, SP)
        h.set_sp_options(linesearch=ls, k0=k0)
        o.set_sp_options(linesearch=ls, k0=k0)
        rh, lh = h.solve(x0, tol=5e-2)
        ro, lo = o.solve(x0, tol=5e-2)
        assert rh["converged"] == ro["converged"] == 1
        if ro["passes"] <= 30:
            assert rh["passes"] == ro["passes"]
            assert (rh["k1_count"], rh["k2_count"], rh["fallback_count"]) == (ro["k1_count"], ro["k2_count"], ro["fallback_count"])
            assert np.array_equal(lh["backtrack_count"], lo["backtrack_count"])
        h.close()


def check_shift_iterate(lib, flat, w):
    """spk_shift_iterate against an independent derivation by tree PATHS: the node reached from the root by the child
    ordinals (c1, .., ct) takes the values of the node reached by (w, c1, .., ct); entries whose source does not exist
    (last stage; non-leaf-keyed entries whose source is a leaf) keep their values."""
    from spock_b200.handle import tree_maps

    h = Handle(flat, CP, lib=lib)
    nx, nu, ny, d, n, nnl = flat.nx, flat.nu, flat.ny, flat.d, flat.n, flat.n_non_leaf
    cf, anc, _, _ = tree_maps(flat.N, d, lib=lib)
    rng = np.random.default_rng(3)
    z0, v0 = rng.standard_normal(h.nz), rng.standard_normal(h.nv)
    h.set_iterate(z0, v0)
    h.shift_iterate(w)
    z, v = h.get_iterate()
    zo, vo = h.z_off, h.v_off

    def path(j):  # child ordinals from the root to node j (1-based)
        p = []
        while j > 1:
            a = int(anc[j - 1])
            p.append(j - int(cf[a - 1]))
            j = a
        return p[::-1]

    def node(p):
        j = 1
        for c in p:
            if j > nnl:
                return None
            j = int(cf[j - 1]) + c
        return j

    ez, ev = z0.copy(), v0.copy()
    for j in range(1, n + 1):
        s = node([w] + path(j))
        if s is None:
            continue
        j0, s0 = j - 1, s - 1
        ez[j0 * nx:(j0 + 1) * nx] = z0[s0 * nx:(s0 + 1) * nx]
        ez[zo[2] + j0] = z0[zo[2] + s0]
        if j > 1:
            ez[zo[3] + j0 - 1] = z0[zo[3] + s0 - 1]
            ev[vo[2] + (j0 - 1) * nx: vo[2] + j0 * nx] = v0[vo[2] + (s0 - 1) * nx: vo[2] + s0 * nx]
            ev[vo[3] + (j0 - 1) * nu: vo[3] + j0 * nu] = v0[vo[3] + (s0 - 1) * nu: vo[3] + s0 * nu]
            ev[vo[4] + j0 - 1], ev[vo[5] + j0 - 1] = v0[vo[4] + s0 - 1], v0[vo[5] + s0 - 1]
        if s <= nnl:
            ez[zo[1] + j0 * nu: zo[1] + (j0 + 1) * nu] = z0[zo[1] + s0 * nu: zo[1] + (s0 + 1) * nu]
            ez[zo[4] + j0 * ny: zo[4] + (j0 + 1) * ny] = z0[zo[4] + s0 * ny: zo[4] + (s0 + 1) * ny]
            ev[vo[0] + j0 * ny: vo[0] + (j0 + 1) * ny] = v0[vo[0] + s0 * ny: vo[0] + (s0 + 1) * ny]
            ev[vo[1] + j0] = v0[vo[1] + s0]
            ev[vo[6] + j0 * (nx + nu): vo[6] + (j0 + 1) * (nx + nu)] = v0[vo[6] + s0 * (nx + nu): vo[6] + (s0 + 1) * (nx + nu)]
    assert np.array_equal(z, ez) and np.array_equal(v, ev)
    h.close()


def check_solve_batch(lib, flats, algs, tol=1e-3):
    """spk_solve_batch (independent models sharing the GPU) == one spk_solve per model: same passes, same iterate; and a
    second, warm-started batch from the persisted state."""
    from spock_b200.handle import solve_batch

    x0s = [0.1 * np.ones(f.nx) * (1 + 0.1 * k) for k, f in enumerate(flats)]
    x1s = [0.9 * x for x in x0s]
    hb = [Handle(f, a, factor_mode=k % 2, lib=lib) for k, (f, a) in enumerate(zip(flats, algs))]
    hs = [Handle(f, a, factor_mode=k % 2, lib=lib) for k, (f, a) in enumerate(zip(flats, algs))]
    for xs in (x0s, x1s):
        rb = solve_batch(hb, xs, tol=tol)
        for k, h in enumerate(hs):
            r, _ = h.solve(xs[k], tol=tol)
            assert rb[k]["converged"] == r["converged"] == 1
            assert rb[k]["passes"] == r["passes"], (k, rb[k]["passes"], r["passes"])
            (zb, vb), (zs, vs) = hb[k].get_iterate(), h.get_iterate()
            assert relerr(zb, zs) < 1e-12 and relerr(vb, vs) < 1e-12
    for h in hb + hs:
        h.close()


def check_offline_device_vs_host(lib, flat, monkeypatch):
    """The offline phase on the device (spk_k_riccati, spk_k_projector: spk_offline_dev.cuh) against the host versions
    (SPK_HOST_OFFLINE=1, spk_offline.cpp) and, through the oracle, against ricatti_offline / the kernel projector of the
    reference (implicit_l.jl:46-97, cpock.jl:71-81,125-128)."""
    out = []
    for host in ("1", "0"):
        monkeypatch.setenv("SPK_HOST_OFFLINE", host)
        h = Handle(flat, CP, lib=lib)
        nodes = sorted({1, 2, max(1, flat.n_non_leaf // 2), flat.n_non_leaf, flat.n_non_leaf + 1, flat.n})
        out.append([h.factors(i) for i in nodes])
        h.close()
    for fh, fd in zip(*out):
        for a, b in zip(fh, fd):
            assert relerr(b, a) < 1e-12 or np.max(np.abs(a)) == 0.0


def custom_mode_flat(N, d, nx, nu, seed=3):
    """A tree whose node modes node_info[j].w are NOT the generator's (j % d) + 1 but a random map (the reference keeps
    node_info in a plain mutable Vector, scenario_tree.jl:73-81, and ricatti_offline / projection_S1! read w node by node,
    implicit_l.jl:71-95,610-680): the Riccati factors then differ from node to node inside a stage."""
    spec = list(get_random_system_specs(N, d, nx, nu))
    w = np.random.default_rng(seed).integers(1, d + 1, size=spec[0].n).astype(np.int32)
    w[0] = 0
    spec[0].w_override = w
    return flatten_problem(*spec)


def check_custom_modes(lib, flat):
    from spock_b200._lib import check, dptr

    nx, nu = flat.nx, flat.nu
    check_operators(lib, flat, 0)
    check_k_steps(lib, flat, CP, 40, factor_mode=0)
    check_k_steps(lib, flat, SP, 10, factor_mode=0)
    h, o = Handle(flat, CP, lib=lib), Oracle(flat, CP)
    for i in sorted({1, 2, 3, flat.n_non_leaf, flat.n_non_leaf + 1, flat.n}):  # per-node factors against ricatti_offline node by node
        ABK, K, Ri = np.zeros(nx * nx), np.zeros(nu * nx), np.zeros(nu * nu)
        check(lib, lib.spk_get_factors(h._h, i, None, dptr(ABK), dptr(K), dptr(Ri), None))
        _, ABKo, Ko, Rc = o.factors(i)
        if i > 1:
            assert relerr(ABK.reshape(nx, nx).T, ABKo) < 1e-12
        if i <= flat.n_non_leaf:
            assert relerr(K.reshape(nx, nu).T, Ko) < 1e-12
            assert relerr(Ri.reshape(nu, nu).T, np.linalg.inv(Rc @ Rc.T)) < 1e-11
    h.close()
    with pytest.raises(Exception, match="SPK_FACTORS_PER_NODE"):  # per-stage tables cannot represent a mode map
        Handle(flat, CP, factor_mode=1, lib=lib)
