"""Problem primitives -- host mirror of the reference's `src/{cost,dynamics,risk_measures,constraints}.jl`.

Only data lives here.  The arithmetic on these objects (cone / box projections, L, L', prox) runs in
the CUDA library; `flatten_problem` turns the objects into the flat arrays `spk_create` takes
(include/spock_b200.h).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Union

import numpy as np

from .scenario_tree import ScenarioTree

# ------------------------------------------------------------------------------------------------ cost


class Diagonal:
    """Stand-in for Julia's `LinearAlgebra.Diagonal` (the 3-argument server_heat generator builds
    `LA.I(nx) * 1e-1`, a Diagonal; examples/server_heat/server_heat.jl:24-37)."""

    def __init__(self, diag):
        self.diag = np.asarray(diag, dtype=np.float64).copy()

    def __eq__(self, other):
        return isinstance(other, Diagonal) and np.array_equal(self.diag, other.diag)


Matrix = Union[np.ndarray, Diagonal]


@dataclass
class Cost:
    """cost.jl:3-7.  Q, R: one matrix per non-root node (index j-1), QN: one per leaf.
    A single matrix instead of a list means "the same matrix for every node"."""

    Q: Union[Matrix, Sequence[Matrix]]
    R: Union[Matrix, Sequence[Matrix]]
    QN: Union[Matrix, Sequence[Matrix]]


@dataclass
class Dynamics:
    """dynamics.jl:1-4.  A[w], B[w] for mode w = 1..d (0-based lists here)."""

    A: List[np.ndarray]
    B: List[np.ndarray]


# ------------------------------------------------------------------------------------------------ cones / risk

CONE_ZEROS, CONE_NONNEG, CONE_NONPOS, CONE_REALS, CONE_SOC = 0, 1, 2, 3, 4


@dataclass(frozen=True)
class _Cone:
    dimension: int


class Zeros(_Cone):
    type_id = CONE_ZEROS


class Nonnegatives(_Cone):
    type_id = CONE_NONNEG


class Nonpositives(_Cone):
    type_id = CONE_NONPOS


class Reals(_Cone):
    type_id = CONE_REALS


class SecondOrderCone(_Cone):
    """MOI order (t, x) (risk_measures.jl:40-71)."""

    type_id = CONE_SOC


@dataclass
class ConvexCone:
    """risk_measures.jl:73-75."""

    subcones: List[_Cone]


@dataclass
class RiskMeasure:
    """risk_measures.jl:79-84: {mu : b - E mu in K, F' ... } in the reference's conic form."""

    E: np.ndarray
    F: np.ndarray
    b: np.ndarray
    K: ConvexCone


class UniformRiskMeasures(Sequence):
    """`Vector{RiskMeasure}` with the same RiskMeasure at every non-leaf node (risk_measures.jl:93-104)
    without materialising n_non_leaf references."""

    def __init__(self, rm: RiskMeasure, count: int):
        self.rm, self.count = rm, int(count)

    def __len__(self):
        return self.count

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self.rm] * len(range(*i.indices(self.count)))
        if not (-self.count <= i < self.count):
            raise IndexError(i)
        return self.rm


def get_uniform_rms_v2(E, F, b, K, d, N) -> UniformRiskMeasures:
    """risk_measures.jl:93-104."""
    return UniformRiskMeasures(
        RiskMeasure(np.asarray(E, float), np.asarray(F, float), np.asarray(b, float), K), (d ** (N - 1) - 1) // (d - 1)
    )


def get_uniform_rms_avar_v2(p, alpha, d, N) -> UniformRiskMeasures:
    """risk_measures.jl:110-119: E = [alpha I; -I; 1'], F = 0, b = [p; 0; 1], K = Nonneg(2d) x Zeros(1)."""
    E = np.vstack([alpha * np.eye(d), -np.eye(d), np.ones((1, d))])
    F = np.zeros((2 * d + 1, d))
    b = np.concatenate([np.asarray(p, float), np.zeros(d), [1.0]])
    return get_uniform_rms_v2(E, F, b, ConvexCone([Nonnegatives(2 * d), Zeros(1)]), d, N)


def get_nonuniform_rms_avar_v2(d, N, rng=None):
    """risk_measures.jl:127-136: a different AV@R at every non-leaf node -- E = [a_i I; -I; 1'] with a_i = rand(),
    b = [p_i; 0; 1] with p_i = rand_probvec2(d).  The reference draws from the unseeded global RNG; here `rng` is a
    numpy Generator (default_rng(0) if None) so that the problem is reproducible."""
    rng = np.random.default_rng(0) if rng is None else rng
    out = []
    for _ in range((d ** (N - 1) - 1) // (d - 1)):
        a = float(rng.random())
        p = rng.random(d)
        p = p / p.sum()
        E = np.vstack([a * np.eye(d), -np.eye(d), np.ones((1, d))])
        out.append(RiskMeasure(E, np.zeros((2 * d + 1, d)), np.concatenate([p, np.zeros(d), [1.0]]),
                               ConvexCone([Nonnegatives(2 * d), Zeros(1)])))
    return out


def get_uniform_rms_tv_v2(p, r, d, N) -> UniformRiskMeasures:
    """risk_measures.jl:138-147 (total-variation ambiguity set)."""
    E = np.vstack([0.5 * np.eye(d), -0.5 * np.eye(d), np.zeros((d, d))])
    F = np.vstack([-np.eye(d), -np.eye(d), np.eye(d)])
    p = np.asarray(p, float)
    b = np.concatenate([0.5 * p, -0.5 * p, np.full(d, float(r))])
    return get_uniform_rms_v2(E, F, b, ConvexCone([Nonnegatives(3 * d)]), d, N)


def get_uniform_rms_l2ball_v2(p, r, d, N) -> UniformRiskMeasures:
    """l2-ball ambiguity set {mu >= 0, 1'mu = 1, ||mu - p|| <= r} in the reference's conic form
    b - E mu in K with K = Nonneg(d) x Zeros(1) x SOC(d+1).
    Not generated by the reference; exercises the SecondOrderCone branch of implicit_l.jl:833-845."""
    E = np.vstack([-np.eye(d), np.ones((1, d)), np.zeros((1, d)), np.eye(d)])
    F = np.zeros((2 * d + 2, d))
    b = np.concatenate([np.zeros(d), [1.0], [float(r)], np.asarray(p, float)])
    return get_uniform_rms_v2(E, F, b, ConvexCone([Nonnegatives(d), Zeros(1), SecondOrderCone(d + 1)]), d, N)


# ------------------------------------------------------------------------------------------------ constraints


@dataclass
class UniformRectangle:
    """constraints.jl:11-22 (same positional fields; the Julia field names nΓ_leaf / nΓ_nonleaf are
    spelled nGamma_* here)."""

    x_min: float
    x_max: float
    u_min: float
    u_max: float
    nGamma_leaf: int
    nGamma_nonleaf: int
    nx: int
    nu: int
    n_leafs: int
    n_nonleafs: int


# ------------------------------------------------------------------------------------------------ flattening


def _sqrt_psd(M: np.ndarray) -> np.ndarray:
    """Principal square root of a symmetric PSD matrix (Julia `sqrt(::Matrix)` at cpock.jl:119-121)."""
    M = np.asarray(M, dtype=np.float64)
    if np.count_nonzero(M - np.diag(np.diagonal(M))) == 0:
        return np.diag(np.sqrt(np.diagonal(M)))
    w, V = np.linalg.eigh(0.5 * (M + M.T))
    return (V * np.sqrt(np.clip(w, 0.0, None))) @ V.T


@dataclass
class FlatProblem:
    """Flat, C-ABI-ready description of one problem (see include/spock_b200.h `spk_problem`).
    Matrices are column-major (Julia layout) inside 1-D float64 arrays."""

    N: int
    d: int
    nx: int
    nu: int
    ny: int
    A: np.ndarray
    B: np.ndarray
    cost_dense: int
    n_sqrtQ: int
    sqrtQ: np.ndarray
    sqrtR: np.ndarray
    n_sqrtQN: int
    sqrtQN: np.ndarray
    E: np.ndarray
    F: np.ndarray
    b: np.ndarray
    cone_type: np.ndarray
    cone_dim: np.ndarray
    x_min: float
    x_max: float
    u_min: float
    u_max: float
    # kept for host-side use (not passed to C)
    E2d: np.ndarray = field(repr=False, default=None)
    F2d: np.ndarray = field(repr=False, default=None)
    w: np.ndarray = field(repr=False, default=None)     # caller-supplied node_info[j].w (1-based, length n), or None
    n_rms: int = 1                                      # 1 = uniform risk measures, n_non_leaf = one (E, F, b) per node
    rms_list: list = field(repr=False, default=None)    # the per-node RiskMeasure objects when n_rms > 1

    @property
    def n(self) -> int:
        return (self.d**self.N - 1) // (self.d - 1)

    @property
    def n_non_leaf(self) -> int:
        return (self.d ** (self.N - 1) - 1) // (self.d - 1)

    @property
    def n_leaf(self) -> int:
        return self.d ** (self.N - 1)

    @property
    def nz(self) -> int:
        n, nnl = self.n, self.n_non_leaf
        return n * self.nx + nnl * self.nu + n + (n - 1) + nnl * self.ny

    @property
    def nv(self) -> int:
        n, nnl, nl = self.n, self.n_non_leaf, self.n_leaf
        return (
            nnl * self.ny + nnl + (n - 1) * (self.nx + self.nu + 2) + nnl * (self.nx + self.nu) + nl * (2 * self.nx + 2)
        )


def _colmajor(M: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(M, dtype=np.float64).T).reshape(-1)


def _flatten_cost_list(mats, count: int, dim: int):
    """Returns (is_dense, n_stored, sqrt matrices as a list) for Q/R/QN given as one matrix or a list."""
    if isinstance(mats, (np.ndarray, Diagonal)):
        mats = [mats]
    mats = list(mats)
    if len(mats) not in (1, count):
        raise ValueError(f"cost list has {len(mats)} entries, expected 1 or {count}")
    first = mats[0]
    uniform = len(mats) == 1 or all(
        (m is first) or (type(m) is type(first) and (m == first if isinstance(m, Diagonal) else np.array_equal(m, first)))
        for m in mats
    )
    if uniform:
        mats = [first]
    dense = not all(isinstance(m, Diagonal) for m in mats)
    out = []
    for m in mats:
        if isinstance(m, Diagonal):
            sq = np.sqrt(m.diag)
            out.append(np.diag(sq) if dense else sq)
        else:
            m = np.asarray(m, dtype=np.float64)
            if m.shape != (dim, dim):
                raise ValueError(f"cost matrix has shape {m.shape}, expected {(dim, dim)}")
            out.append(_sqrt_psd(m))
    return dense, out


def flatten_problem(scen_tree: ScenarioTree, cost: Cost, dynamics: Dynamics, rms, constraints: UniformRectangle) -> FlatProblem:
    """What `build_cpock` / `build_spock` read out of their arguments (cpock.jl:1-60, 119-121)."""
    nx, nu = np.asarray(dynamics.B[0]).shape  # cpock.jl:12
    d = len(dynamics.A)
    N = scen_tree.N
    if d != scen_tree.d:
        raise ValueError("number of dynamics modes must equal the branching factor")
    n, nl = scen_tree.n, scen_tree.n_leaf_nodes
    rm = rms[0]
    ny = len(rm.b)  # implicit_l.jl:20: one ny for all nodes, like the reference
    uniform = isinstance(rms, UniformRiskMeasures) or all(
        r is rm or (np.array_equal(r.E, rm.E) and np.array_equal(r.F, rm.F) and np.array_equal(r.b, rm.b)) for r in rms)
    if not uniform:
        if len(rms) != scen_tree.n_non_leaf_nodes:
            raise ValueError("one risk measure per non-leaf node expected")
        for r in rms:
            if len(r.b) != ny or [(c.type_id, c.dimension) for c in r.K.subcones] != [(c.type_id, c.dimension) for c in rm.K.subcones]:
                raise NotImplementedError("risk measures must share ny and the sub-cone list (implicit_l.jl:14,19)")
    E = np.asarray(rm.E, float)
    F = np.asarray(rm.F, float)
    if E.shape != (ny, d) or F.shape[0] != ny:
        raise ValueError("risk measure E/F have inconsistent shapes")
    if F.shape[1] != d:
        raise ValueError("F must have d columns (cpock.jl:76-78)")
    if sum(c.dimension for c in rm.K.subcones) != ny:
        raise ValueError("sub-cone dimensions must add up to length(b)")
    dq, sQ = _flatten_cost_list(cost.Q, n - 1, nx)
    dr, sR = _flatten_cost_list(cost.R, n - 1, nu)
    dqn, sQN = _flatten_cost_list(cost.QN, nl, nx)
    dense = dq or dr or dqn
    if len(sQ) != len(sR):  # one uniform, one per node: replicate the uniform one
        cnt = max(len(sQ), len(sR))
        sQ = sQ * cnt if len(sQ) == 1 else sQ
        sR = sR * cnt if len(sR) == 1 else sR

    def pack(lst, was_dense, dim):
        if dense:
            return np.concatenate([_colmajor(m if was_dense else np.diag(m)) for m in lst])
        return np.concatenate([np.asarray(m, float).reshape(-1) for m in lst])

    return FlatProblem(
        N=N, d=d, nx=nx, nu=nu, ny=ny,
        A=np.concatenate([_colmajor(a) for a in dynamics.A]),
        B=np.concatenate([_colmajor(b) for b in dynamics.B]),
        cost_dense=int(dense),
        n_sqrtQ=len(sQ), sqrtQ=pack(sQ, dq, nx), sqrtR=pack(sR, dr, nu),
        n_sqrtQN=len(sQN), sqrtQN=pack(sQN, dqn, nx),
        E=_colmajor(E) if uniform else np.concatenate([_colmajor(np.asarray(r.E, float)) for r in rms]),
        F=_colmajor(F) if uniform else np.concatenate([_colmajor(np.asarray(r.F, float)) for r in rms]),
        b=np.asarray(rm.b, float).copy() if uniform else np.concatenate([np.asarray(r.b, float) for r in rms]),
        w=None if getattr(scen_tree, "w_override", None) is None else scen_tree.w_array(),
        n_rms=1 if uniform else len(rms),
        rms_list=None if uniform else list(rms),
        cone_type=np.array([c.type_id for c in rm.K.subcones], dtype=np.int32),
        cone_dim=np.array([c.dimension for c in rm.K.subcones], dtype=np.int32),
        x_min=float(constraints.x_min), x_max=float(constraints.x_max),
        u_min=float(constraints.u_min), u_max=float(constraints.u_max),
        E2d=E, F2d=F,
    )
