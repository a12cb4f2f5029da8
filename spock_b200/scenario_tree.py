"""Scenario tree -- host mirror of the reference's `src/scenario_tree.jl`.

The index maps are part of the parity contract (bit-exact with the reference):
1-based BFS numbering, `children(i) = d(i-1)+2 .. d(i-1)+d+1`, `anc(j) = (j-2)//d + 1`,
mode `w(j) = (j % d) + 1` (scenario_tree.jl:61-109).  All values stored here are the reference's
1-based node numbers; nothing is materialised per node unless asked for, so a 22M-node tree
costs a few integers.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np


@dataclass(frozen=True)
class ScenarioTreeNodeInfo:
    """scenario_tree.jl:25-31 (1-based index lists, None where the reference stores `nothing`)."""

    x: Optional[List[int]]
    u: Optional[List[int]]
    w: Optional[int]
    s: Optional[int]
    tau: Optional[int]


class _ChildMapping:
    """Read-only `Dict{Int, Vector{Int}}` view: child_mapping[i] for non-leaf i (scenario_tree.jl:82-88)."""

    def __init__(self, tree: "ScenarioTree"):
        self._t = tree

    def __getitem__(self, i: int) -> List[int]:
        if not (1 <= i <= self._t.n_non_leaf_nodes):
            raise KeyError(i)
        first = self._t.d * (i - 1) + 2
        return list(range(first, first + self._t.d))

    def __len__(self) -> int:
        return self._t.n_non_leaf_nodes

    def __contains__(self, i: int) -> bool:
        return 1 <= i <= self._t.n_non_leaf_nodes

    def keys(self):
        return range(1, self._t.n_non_leaf_nodes + 1)

    def items(self):
        return ((i, self[i]) for i in self.keys())


class _AncMapping:
    """Read-only `Dict{Int, Int}` view: anc_mapping[j] for j >= 2 (scenario_tree.jl:90-95)."""

    def __init__(self, tree: "ScenarioTree"):
        self._t = tree

    def __getitem__(self, j: int) -> int:
        if not (2 <= j <= self._t.n):
            raise KeyError(j)
        return (j - 2) // self._t.d + 1

    def __len__(self) -> int:
        return self._t.n - 1

    def __contains__(self, j: int) -> bool:
        return 2 <= j <= self._t.n


class _NodeInfo:
    """Lazy `Vector{ScenarioTreeNodeInfo}` (scenario_tree.jl:73-81); node_info[i] with 1-based i."""

    def __init__(self, tree: "ScenarioTree"):
        self._t = tree

    def __len__(self) -> int:
        return self._t.n

    def __getitem__(self, i: int) -> ScenarioTreeNodeInfo:
        t = self._t
        if not (1 <= i <= t.n):
            raise IndexError(i)
        return ScenarioTreeNodeInfo(
            x=list(range((i - 1) * t.nx + 1, i * t.nx + 1)),
            u=list(range((i - 1) * t.nu + 1, i * t.nu + 1)) if i <= t.n_non_leaf_nodes else None,
            w=(int(t.w_override[i - 1]) if t.w_override is not None else (i % t.d) + 1) if i > 1 else None,
            s=i,
            tau=i if i > 1 else None,
        )


@dataclass
class ScenarioTree:
    """scenario_tree.jl:44-55, plus the generator arguments (d, nx, nu) the Julia struct does not keep."""

    N: int
    d: int
    nx: int
    nu: int
    n: int
    n_non_leaf_nodes: int
    n_leaf_nodes: int
    leaf_node_min_index: int
    leaf_node_max_index: int
    min_index_per_timestep: List[int]
    # node_info[j].w supplied by the caller instead of the generator's (j % d) + 1 (the reference's node_info vector is a
    # plain mutable Vector, scenario_tree.jl:73-81): 1-based modes, length n, entry 0 (the root) ignored
    w_override: Optional[np.ndarray] = None

    @property
    def child_mapping(self) -> _ChildMapping:
        return _ChildMapping(self)

    @property
    def anc_mapping(self) -> _AncMapping:
        return _AncMapping(self)

    @property
    def node_info(self) -> _NodeInfo:
        return _NodeInfo(self)

    # ---- vectorised maps (reference 1-based values) used by tests and by the flattening code
    def child_first_array(self) -> np.ndarray:
        i = np.arange(1, self.n_non_leaf_nodes + 1, dtype=np.int64)
        return self.d * (i - 1) + 2

    def anc_array(self) -> np.ndarray:
        j = np.arange(1, self.n + 1, dtype=np.int64)
        a = (j - 2) // self.d + 1
        a[0] = 0
        return a

    def w_array(self) -> np.ndarray:
        if self.w_override is not None:
            w = np.array(self.w_override, dtype=np.int64)
        else:
            j = np.arange(1, self.n + 1, dtype=np.int64)
            w = (j % self.d) + 1
        w[0] = 0
        return w.astype(np.int32)


def generate_scenario_tree_uniform_branching_factor_v2(N: int, d: int, nx: int, nu: int) -> ScenarioTree:
    """scenario_tree.jl:61-109."""
    if d <= 1:
        raise ValueError(f"Branching factor d must be larger than 1, but is {d}.")
    n_total = (d**N - 1) // (d - 1)
    n_leafs = d ** (N - 1)
    n_non_leafs = (d ** (N - 1) - 1) // (d - 1)
    return ScenarioTree(
        N=N,
        d=d,
        nx=nx,
        nu=nu,
        n=n_total,
        n_non_leaf_nodes=n_non_leafs,
        n_leaf_nodes=n_leafs,
        leaf_node_min_index=n_non_leafs + 1,
        leaf_node_max_index=n_total,
        min_index_per_timestep=[1] + [1 + (d**i - 1) // (d - 1) for i in range(1, N)],
    )
