"""The oracle is pinned against the reference's committed traces
(examples/server_heat/output/residuals.tikz -> tests/golden/, see tests/golden/make_golden.py)."""
import numpy as np

from oracle.oracle import CP, SP, Oracle
from tests.helpers import flat_of, golden_cp, golden_sp


def test_cp_trace_all_3159_iterations():
    gold = golden_cp()
    o = Oracle(flat_of("heat_N7_d2_nx5"), CP)
    res, log = o.solve(0.1 * np.ones(5), tol=1e-5, verbose_log=True)
    assert res["passes"] == 3159 == len(gold) and res["converged"] == 1 and res["printed_iter"] == 3158
    assert np.max(np.abs(log["xi"] - gold) / gold) < 1e-9  # measured: 2.9e-11
    assert (o.nz, o.nv, o.n) == (1518, 3288, 127)


def test_sp_trace_first_iterations_and_branch_pattern():
    xi, bc = golden_sp()
    o = Oracle(flat_of("heat_N7_d2_nx5"), SP)
    res, log = o.solve(0.1 * np.ones(5), tol=1e-5, verbose_log=True)
    # SuperMann + Anderson amplifies rounding chaotically: pinned for the first 30 passes
    assert np.max(np.abs(log["xi"][:30] - xi[:30]) / xi[:30]) < 1e-9
    assert np.max(np.abs(log["xi"][:45] - xi[:45]) / xi[:45]) < 1e-8
    assert np.array_equal(log["backtrack_count"][:100], bc[:100])
    assert set(np.unique(log["backtrack_count"])) <= {0, 9}
    assert res["converged"] == 1 and log["xi"][-1] <= 1e-5
