"""Tree index maps are bit-exact with the reference (scenario_tree.jl:61-109): python mirror == C oracle ==
library, plus the known answers pinned by the reference's own tests (test/model_dynamics/implicit_L.jl:53-65)."""
import numpy as np
import pytest

from oracle.oracle import CP, Oracle
from spock_b200 import _lib, generate_scenario_tree_uniform_branching_factor_v2
from spock_b200.handle import tree_maps
from tests.helpers import flat_of


def _lib_nogpu():
    from spock_b200 import build
    build.build()
    return _lib.load()


@pytest.mark.parametrize("N,d", [(3, 2), (7, 2), (4, 3), (5, 4), (2, 5), (13, 2)])
def test_maps_agree(N, d):
    t = generate_scenario_tree_uniform_branching_factor_v2(N, d, 2, 1)
    cf, anc, w, mi = tree_maps(N, d, lib=_lib_nogpu())  # pure host function of the CUDA library: no device needed
    assert np.array_equal(cf, t.child_first_array())
    assert np.array_equal(anc, t.anc_array())
    assert np.array_equal(w, t.w_array())
    assert list(mi) == t.min_index_per_timestep
    assert t.n == (d**N - 1) // (d - 1) and t.n_leaf_nodes == d ** (N - 1)
    assert t.leaf_node_min_index == t.n_non_leaf_nodes + 1 and t.leaf_node_max_index == t.n


@pytest.mark.parametrize("name", ["fixture_N3_d2", "heat_N4_d3_nx3"])
def test_maps_agree_with_oracle(name):
    flat = flat_of(name)
    o = Oracle(flat, CP)
    cf, anc, w, mi = o.tree_maps()
    cf2, anc2, w2, mi2 = tree_maps(flat.N, flat.d, lib=_lib_nogpu())
    assert np.array_equal(cf, cf2) and np.array_equal(anc, anc2) and np.array_equal(w, w2) and np.array_equal(mi, mi2)


def test_reference_known_answers():
    # 7-node fixture: children and modes used by test/model_dynamics/implicit_L.jl:60-65
    t = generate_scenario_tree_uniform_branching_factor_v2(3, 2, 2, 1)
    assert t.child_mapping[1] == [2, 3] and t.child_mapping[2] == [4, 5] and t.child_mapping[3] == [6, 7]
    assert [t.node_info[j].w for j in range(2, 8)] == [1, 2, 1, 2, 1, 2]
    assert t.node_info[1].w is None and t.node_info[4].u is None and t.node_info[3].u == [3]
    assert t.node_info[2].x == [3, 4] and t.anc_mapping[7] == 3 and t.min_index_per_timestep == [1, 2, 4]
    # d = 3: children of the root get modes 3, 1, 2 (w = (j % d) + 1)
    t3 = generate_scenario_tree_uniform_branching_factor_v2(3, 3, 1, 1)
    assert [t3.node_info[j].w for j in (2, 3, 4)] == [3, 1, 2]
    with pytest.raises(ValueError):
        generate_scenario_tree_uniform_branching_factor_v2(3, 1, 1, 1)  # scenario_tree.jl:62-64


def test_z_layout_of_fixture():
    # test/model_dynamics/implicit_L.jl:53-57: x=z[1:14], u=z[15:17], s=z[18:24], tau=z[25:30], y=z[31:45]
    flat = flat_of("fixture_N3_d2")
    o = Oracle(flat, CP)
    zo, vo = o.offsets()
    assert list(zo) == [0, 14, 17, 24, 30] and o.nz == 45 and o.nv == 81
    assert flat.nz == 45 and flat.nv == 81
